"""UNet-A: drop-in for the reference's networks/unet_model.py (the model train.py / train_mnms.py /
test.py build: ``UNet(n_channels=, n_classes=)``, unet_model.py:6-39).

Same constructor, attributes, sub-module names (=> identical state_dict keys and seeded init),
``forward(x, feature=False)``.  The forward is one program of sm_100a kernels:

  inc/down*: conv3x3 (tcgen05 implicit GEMM, BN statistics in the epilogue) -> BN finalize ->
             fused BN-apply+ReLU that writes the skip tensor directly into the decoder's concat
             buffer and the 2x2-max-pooled tensor for the next level;
  up*:       ConvTranspose2d as 4 tcgen05 GEMMs TMA-storing into the other half of that concat
             buffer (no F.pad clone, no torch.cat copy: unet_parts.py:62-67);
  outc:      1x1 head streamed by a CUDA-core kernel to fp32 NCHW logits.
"""
import torch.nn as nn

from ustrun import engine as E
from ustrun.bridge import Feature, run_program

from .unet_parts import DoubleConv, Down, OutConv, Up


class UNet(nn.Module):
    def __init__(self, n_channels, n_classes, bilinear=False, norm='bn', num_domains=None):
        """``norm`` / ``num_domains`` are an extension (upstream: plain BatchNorm only, unet_model.py:7): 'dsbn' puts the
        reference's DomainSpecificBatchNorm2d (networks/dsbn.py) behind every conv, ``forward(x, domain_label=...)``."""
        super(UNet, self).__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        factor = 2 if bilinear else 1
        kw = dict(norm=norm, num_domains=num_domains)
        self.inc = DoubleConv(n_channels, 64, **kw)
        self.down1 = Down(64, 128, **kw)
        self.down2 = Down(128, 256, **kw)
        self.down3 = Down(256, 512, **kw)
        self.down4 = Down(512, 1024 // factor, **kw)
        self.up1 = Up(1024, 512 // factor, bilinear, **kw)
        self.up2 = Up(512, 256 // factor, bilinear, **kw)
        self.up3 = Up(256, 128 // factor, bilinear, **kw)
        self.up4 = Up(128, 64, bilinear, **kw)
        self.outc = OutConv(64, n_classes)

    def program(self, ctx, a, feature=False, domain_label=None):
        """Engine program for one forward (a: NHWC input activation)."""
        if a.H % 16 or a.W % 16:
            raise ValueError("UNet input height/width must be divisible by 16, got {}x{}".format(a.H, a.W))
        B, H, W = a.B, a.H, a.W
        enc = [self.inc, self.down1.dc, self.down2.dc, self.down3.dc, self.down4.dc]
        skip_c = [64, 128, 256, 512]
        # concat buffers of up4..up1: [skip | up-sampled]; both halves have skip_c[i] channels
        # (ConvTranspose halves the channels of the level below; bilinear=True halves them in down4/DoubleConv)
        up_c = skip_c
        cats = [E.Act.new(B, H >> i, W >> i, skip_c[i] + up_c[i], dtype=a.t.dtype, device=a.t.device) for i in range(4)]
        h = a
        for i in range(4):
            _, h = enc[i].run(ctx, h, out=cats[i].view(0, skip_c[i]), pool=True, domain_label=domain_label)
        h, _ = enc[4].run(ctx, h, domain_label=domain_label)
        for i, up in zip((3, 2, 1, 0), (self.up1, self.up2, self.up3, self.up4)):
            h = up.run(ctx, h, cats[i], skip_c[i], domain_label=domain_label)
        head = self.outc.run(ctx, h)
        return (head, Feature(h)) if feature else (head,)

    def forward(self, x, feature=False, domain_label=None):
        out = run_program(self, lambda ctx, a: self.program(ctx, a, feature, domain_label), x)
        return (out[0], out[1]) if feature else out[0]
