"""Building blocks of UNet-A (drop-in for the reference's networks/unet_parts.py).

Each class keeps the reference's constructor signature, sub-module names and registration order
(so ``state_dict()`` keys, ``parameters()`` order and seeded initialisation are identical:
unet_parts.py:8-25 DoubleConv, :28-39 Down, :42-68 Up, :71-77 OutConv), but the sub-modules are
only parameter containers: the arithmetic is done by ``run()`` programs made of engine ops, i.e.
by the sm_100a kernels (conv + fused BN statistics, BN-apply+ReLU(+MaxPool), ConvTranspose written
straight into the concat buffer).  Calling a block on its own (``block(x)``) executes the same
program through a small autograd bridge.
"""
import torch
import torch.nn as nn

from ustrun import engine as E
from ustrun import _lib as L
from ustrun.bridge import run_program


def _norm_layer(channels, norm, num_domains):
    if norm == 'bn':
        return nn.BatchNorm2d(channels)
    if norm == 'dsbn':
        from networks.dsbn import DomainSpecificBatchNorm2d
        return DomainSpecificBatchNorm2d(channels, num_domains=num_domains)
    raise ValueError('Normalization type {} is not supporter'.format(norm))


def _select_bn(m, domain_label):
    if isinstance(m, nn.BatchNorm2d):
        return m
    if domain_label is None:
        raise TypeError("forward() missing 1 required positional argument: 'domain_label'")
    return m.select(domain_label)


class DoubleConv(nn.Module):
    """(conv3x3 no-bias -> BatchNorm -> ReLU) x 2.  ``norm='dsbn'`` (not upstream: BASELINE.json configs[2], "DSBN over 3
    domains" on the tensor-bound model) replaces both BatchNorm2d by the reference's DomainSpecificBatchNorm2d."""

    def __init__(self, in_channels, out_channels, mid_channels=None, norm='bn', num_domains=None):
        super().__init__()
        mid_channels = mid_channels or out_channels
        layers = [nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), _norm_layer(mid_channels, norm, num_domains),
                  nn.ReLU(inplace=True),
                  nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), _norm_layer(out_channels, norm, num_domains),
                  nn.ReLU(inplace=True)]
        self.double_conv = nn.Sequential(*layers)
        self._pk = [E.PackedConv(), E.PackedConv()]
        self.out_channels = out_channels

    def run(self, ctx, x, out=None, pool=False, domain_label=None):
        s = self.double_conv
        mid, _ = E.conv_bn_act(ctx, x, s[0], _select_bn(s[1], domain_label), L.ACT_RELU, packed=self._pk[0])
        return E.conv_bn_act(ctx, mid, s[3], _select_bn(s[4], domain_label), L.ACT_RELU, out=out, pool=pool, packed=self._pk[1])

    def forward(self, x, domain_label=None):
        return run_program(self, lambda ctx, a: (self.run(ctx, a, domain_label=domain_label)[0],), x)[0]


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv.  In a full UNet the pooling is fused into the producer's
    BN-apply kernel; stand-alone the block pools its own input."""

    def __init__(self, in_channels, out_channels, norm='bn', num_domains=None):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels, norm=norm, num_domains=num_domains))

    @property
    def dc(self):
        return self.maxpool_conv[1]

    def forward(self, x):
        raise NotImplementedError("Down is executed as part of UNet.forward (the 2x2 max-pool is fused into the "
                                  "previous block's BN-apply kernel); it has no stand-alone sm_100a program")


class Up(nn.Module):
    """Up-scaling (ConvTranspose2d k2 s2, or bilinear align_corners=True) + concat + DoubleConv."""

    def __init__(self, in_channels, out_channels, bilinear=True, norm='bn', num_domains=None):
        super().__init__()
        self.bilinear = bilinear
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2, norm=norm, num_domains=num_domains)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels, norm=norm, num_domains=num_domains)
        self._pk = E.PackedConv()

    def run(self, ctx, x1, cat, skip_channels, domain_label=None):
        """x1: low-res input; cat: concat buffer whose first ``skip_channels`` channels already hold
        the skip tensor (torch.cat([x2, x1], 1) of unet_parts.py:67 without the copy)."""
        if x1.H * 2 != cat.H or x1.W * 2 != cat.W:
            raise ValueError("UNet input height/width must be divisible by 16 (F.pad path of unet_parts.py:59-63 is not implemented)")
        up_view = cat.view(skip_channels, cat.C - skip_channels)
        if self.bilinear:
            E.upsample2x(ctx, x1, True, out=up_view)
        else:
            E.conv_transpose2x2(ctx, x1, self.up, up_view, self._pk)
        cat.needs_grad = ctx.need_grad
        return self.conv.run(ctx, cat, domain_label=domain_label)[0]

    def forward(self, x1, x2):
        raise NotImplementedError("Up is executed as part of UNet.forward (it writes into the shared concat buffer)")


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super(OutConv, self).__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self._pk = E.PackedConv()

    def run(self, ctx, x):
        return E.head_conv(ctx, x, self.conv, self._pk)

    def forward(self, x):
        raise NotImplementedError("OutConv is executed as part of UNet.forward")
