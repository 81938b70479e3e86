"""UNet-B family: drop-in for the reference's networks/unet.py (n=16 base width, three convs per
block with conv bias, bilinear up-sampling, optional domain-specific BatchNorm).

Constructor signatures, sub-module names and registration order follow the reference (ConvD
unet.py:32-72, ConvU :75-117, ConvU_Rec :120-165, Unet2D :168-203, Unet2D_MT :206-245, Encoder
:248-271, Decoder :273-296, UNet :298-334, Rec_Decoder :339-362, Unet2D_DS :365-418, Unet2D_MS
:421-470) so state_dict keys and seeded initialisation (kaiming_normal fan_out, BN 1/0) match.
The sub-modules are parameter holders; ``run()`` programs execute on the sm_100a kernels:
the conv bias in front of a train-mode BatchNorm is never added (it cancels; only the running
mean sees it), BN-apply/activation/max-pool are one kernel, torch.cat is replaced by producers
writing into channel slices of one NHWC buffer.  Only 'bn' and 'dsbn' normalisation have
kernels; 'gn'/'in' modules can be constructed (checkpoint parity) but raise on forward.

Extension over upstream: ConvD/ConvU accept ``num_domains`` and the networks accept
``domain_label=`` so that ``UNet(norm='dsbn', num_domains=k)`` is constructible; with the default
arguments behaviour equals upstream (where norm='dsbn' raises TypeError at construction).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ustrun import engine as E
from ustrun import _lib as L
from ustrun.bridge import Feature, run_program

from networks.dsbn import DomainSpecificBatchNorm2d


def count_params(model):
    return sum(p.numel() for p in model.parameters()) / 1e6


def normalization(planes, norm='gn', num_domains=None):
    if norm == 'bn':
        m = nn.BatchNorm2d(planes)
    elif norm == 'gn':
        m = nn.GroupNorm(1, planes)
    elif norm == 'in':
        m = nn.InstanceNorm2d(planes)
    elif norm == 'dsbn':
        m = DomainSpecificBatchNorm2d(planes, num_domains=num_domains)
    else:
        raise ValueError('Normalization type {} is not supporter'.format(norm))
    return m


def _act_code(module):
    return L.ACT_RELU if isinstance(module, nn.ReLU) else L.ACT_LEAKY


def _make_activation(activation):
    return nn.ReLU(inplace=True) if activation == 'relu' else nn.LeakyReLU(0.01, inplace=True)


def _pick_bn(norm_module, domain_label):
    """Resolve the nn.BatchNorm2d a block uses for this forward."""
    if isinstance(norm_module, nn.BatchNorm2d):
        return norm_module
    if isinstance(norm_module, DomainSpecificBatchNorm2d):
        if domain_label is None:          # upstream calls dsbn(x) without the label -> TypeError
            raise TypeError("forward() missing 1 required positional argument: 'domain_label'")
        return norm_module.select(domain_label)
    raise NotImplementedError("only norm='bn' and norm='dsbn' have sm_100a kernels (got {}); there is no "
                              "PyTorch fallback".format(type(norm_module).__name__))


def _init_weights(module, activation):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity=activation)
        elif isinstance(m, nn.BatchNorm2d) or isinstance(m, nn.GroupNorm):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class ConvD(nn.Module):
    def __init__(self, inplanes, planes, norm='bn', first=False, activation='relu', num_domains=None):
        super(ConvD, self).__init__()
        self.first = first
        self.conv1 = nn.Conv2d(inplanes, planes, 3, 1, 1, bias=True)
        self.bn1 = normalization(planes, norm, num_domains)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=True)
        self.bn2 = normalization(planes, norm, num_domains)
        self.conv3 = nn.Conv2d(planes, planes, 3, 1, 1, bias=True)
        self.bn3 = normalization(planes, norm, num_domains)
        self.maxpool2D = nn.MaxPool2d(kernel_size=2)
        self.activation = _make_activation(activation)
        self._pk = [E.PackedConv() for _ in range(3)]

    def run(self, ctx, x, out=None, pool=False, domain_label=None):
        """x is this block's input AFTER the max-pool (the pool of unet.py:55-56 is fused into the
        previous block's last kernel, which is what ``pool=True`` requests for THIS block's output).
        Note: no activation after bn1 (unet.py:58-60)."""
        act = _act_code(self.activation)
        t, _ = E.conv_bn_act(ctx, x, self.conv1, _pick_bn(self.bn1, domain_label), L.ACT_NONE, packed=self._pk[0])
        t, _ = E.conv_bn_act(ctx, t, self.conv2, _pick_bn(self.bn2, domain_label), act, packed=self._pk[1])
        return E.conv_bn_act(ctx, t, self.conv3, _pick_bn(self.bn3, domain_label), act, out=out, pool=pool, packed=self._pk[2])

    def forward(self, x):
        raise NotImplementedError("ConvD is executed as part of a network forward (fused max-pool / concat buffers)")


class ConvU(nn.Module):
    def __init__(self, planes, norm='bn', first=False, activation='relu', num_domains=None):
        super(ConvU, self).__init__()
        self.first = first
        if not self.first:
            self.conv1 = nn.Conv2d(2 * planes, planes, 3, 1, 1, bias=True)
            self.bn1 = normalization(planes, norm, num_domains)
        self.pool = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
        self.conv2 = nn.Conv2d(planes, planes // 2, 1, 1, 0, bias=True)
        self.bn2 = normalization(planes // 2, norm, num_domains)
        self.conv3 = nn.Conv2d(planes, planes, 3, 1, 1, bias=True)
        self.bn3 = normalization(planes, norm, num_domains)
        self.activation = _make_activation(activation)
        self.planes = planes
        self._pk = [E.PackedConv() for _ in range(3)]

    def run(self, ctx, x, cat, domain_label=None):
        """cat: [B,2H,2W,planes] buffer whose first planes/2 channels hold ``prev`` (unet.py:110)."""
        act = _act_code(self.activation)
        if not self.first:
            x, _ = E.conv_bn_act(ctx, x, self.conv1, _pick_bn(self.bn1, domain_label), act, packed=self._pk[0])
        up = E.upsample2x(ctx, x, False)
        half = self.planes // 2
        E.conv_bn_act(ctx, up, self.conv2, _pick_bn(self.bn2, domain_label), act, out=cat.view(half, half), packed=self._pk[1])
        cat.needs_grad = ctx.need_grad
        return E.conv_bn_act(ctx, cat, self.conv3, _pick_bn(self.bn3, domain_label), act, packed=self._pk[2])[0]

    def forward(self, x, prev):
        raise NotImplementedError("ConvU is executed as part of a network forward (shared concat buffer)")


class ConvU_Rec(nn.Module):
    def __init__(self, planes, norm='bn', activation='relu', num_domains=None):
        super(ConvU_Rec, self).__init__()
        self.conv1 = nn.Conv2d(planes, planes // 2, 3, 1, 1, bias=True)
        self.bn1 = normalization(planes // 2, norm, num_domains)
        self.pool = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
        self.conv2 = nn.Conv2d(planes // 2, planes // 2, 1, 1, 0, bias=True)
        self.bn2 = normalization(planes // 2, norm, num_domains)
        self.conv3 = nn.Conv2d(planes // 2, planes // 2, 3, 1, 1, bias=True)
        self.bn3 = normalization(planes // 2, norm, num_domains)
        self.activation = _make_activation(activation)
        self._pk = [E.PackedConv() for _ in range(3)]

    def run(self, ctx, x, domain_label=None):
        act = _act_code(self.activation)
        x, _ = E.conv_bn_act(ctx, x, self.conv1, _pick_bn(self.bn1, domain_label), act, packed=self._pk[0])
        y = E.upsample2x(ctx, x, False)
        y, _ = E.conv_bn_act(ctx, y, self.conv2, _pick_bn(self.bn2, domain_label), act, packed=self._pk[1])
        return E.conv_bn_act(ctx, y, self.conv3, _pick_bn(self.bn3, domain_label), act, packed=self._pk[2])[0]

    def forward(self, x, domain_label=None):
        return run_program(self, lambda ctx, a: (self.run(ctx, a, domain_label),), x)[0]


class _Head(nn.Module):
    """Mixin helpers shared by the whole-network classes."""

    def _encode(self, ctx, a, domain_label=None):
        """convd1..5 -> (x5, [cat4..cat1]) with x1..x4 already sitting in their concat buffers."""
        if a.H % 16 or a.W % 16:
            raise ValueError("UNet input height/width must be divisible by 16, got {}x{}".format(a.H, a.W))
        blocks = [self.convd1, self.convd2, self.convd3, self.convd4, self.convd5]
        cats, h = [], a
        for i in range(4):
            c = blocks[i].conv3.out_channels
            cat = E.Act.new(a.B, a.H >> i, a.W >> i, 2 * c, dtype=a.t.dtype, device=a.t.device)
            _, h = blocks[i].run(ctx, h, out=cat.view(0, c), pool=True, domain_label=domain_label)
            cats.append(cat)
        x5, _ = blocks[4].run(ctx, h, domain_label=domain_label)
        return x5, cats

    def _decode(self, ctx, x5, cats, domain_label=None):
        y = self.convu4.run(ctx, x5, cats[3], domain_label)
        y = self.convu3.run(ctx, y, cats[2], domain_label)
        y = self.convu2.run(ctx, y, cats[1], domain_label)
        return self.convu1.run(ctx, y, cats[0], domain_label)


def _build_encoder(self, c, n, norm, activation, num_domains=None):
    self.convd1 = ConvD(c, n, norm, first=True, activation=activation, num_domains=num_domains)
    self.convd2 = ConvD(n, 2 * n, norm, activation=activation, num_domains=num_domains)
    self.convd3 = ConvD(2 * n, 4 * n, norm, activation=activation, num_domains=num_domains)
    self.convd4 = ConvD(4 * n, 8 * n, norm, activation=activation, num_domains=num_domains)
    self.convd5 = ConvD(8 * n, 16 * n, norm, activation=activation, num_domains=num_domains)


def _build_decoder(self, n, norm, activation, num_domains=None):
    self.convu4 = ConvU(16 * n, norm, first=True, activation=activation, num_domains=num_domains)
    self.convu3 = ConvU(8 * n, norm, activation=activation, num_domains=num_domains)
    self.convu2 = ConvU(4 * n, norm, activation=activation, num_domains=num_domains)
    self.convu1 = ConvU(2 * n, norm, activation=activation, num_domains=num_domains)


class Unet2D(_Head):
    def __init__(self, c=3, n=16, norm='bn', num_classes=2, activation='relu'):
        super(Unet2D, self).__init__()
        _build_encoder(self, c, n, norm, activation)
        _build_decoder(self, n, norm, activation)
        self.seg1 = nn.Conv2d(2 * n, num_classes, 3, padding=1)
        self._pk_seg1 = E.PackedConv()
        _init_weights(self, activation)

    def program(self, ctx, a):
        x5, cats = self._encode(ctx, a)
        return (E.head_conv(ctx, self._decode(ctx, x5, cats), self.seg1, self._pk_seg1),)

    def forward(self, x):
        return run_program(self, self.program, x)[0]


class Unet2D_MT(_Head):
    def __init__(self, c=3, n=16, norm='bn', num_classes=2, activation='relu'):
        super(Unet2D_MT, self).__init__()
        _build_encoder(self, c, n, norm, activation)
        _build_decoder(self, n, norm, activation)
        self.seg1 = nn.Conv2d(2 * n, num_classes, 3, padding=1)
        self.rec1 = nn.Conv2d(2 * n, c, 3, padding=1)
        self._pk_seg1, self._pk_rec1 = E.PackedConv(), E.PackedConv()
        _init_weights(self, activation)

    def forward(self, x, is_rec=False):
        def program(ctx, a):
            x5, cats = self._encode(ctx, a)
            y = self._decode(ctx, x5, cats)
            return (E.head_conv(ctx, y, self.rec1, self._pk_rec1) if is_rec else E.head_conv(ctx, y, self.seg1, self._pk_seg1),)
        return run_program(self, program, x)[0]


class Encoder(_Head):
    def __init__(self, c=3, n=16, norm='bn', activation='relu'):
        super(Encoder, self).__init__()
        _build_encoder(self, c, n, norm, activation)
        _init_weights(self, activation)

    def forward(self, x):
        """Returns [x1..x5] as NCHW fp32 tensors.  Only x5 is differentiable when the encoder is
        used on its own (the skips' gradient enters through the fused decoder in a full network)."""
        def program(ctx, a):
            x5, cats = self._encode(ctx, a)
            return tuple(Feature(cat.view(0, cat.C // 2)) for cat in cats) + (x5,)
        return list(run_program(self, program, x))


class Decoder(nn.Module):
    def __init__(self, n=16, num_classes=2, norm='bn', activation='relu'):
        super(Decoder, self).__init__()
        _build_decoder(self, n, norm, activation)
        self.out1 = nn.Conv2d(2 * n, num_classes, 3, padding=1)
        _init_weights(self, activation)

    def forward(self, feats):
        raise NotImplementedError("Decoder is executed inside UNet.forward (it consumes the encoder's concat buffers); "
                                  "use networks.unet.UNet for the fused sm_100a program")


class UNet(_Head):
    def __init__(self, n_channels=3, n_classes=2, n=16, norm='bn', activation='relu', num_domains=None):
        super(UNet, self).__init__()
        _build_encoder(self, n_channels, n, norm, activation, num_domains)
        _build_decoder(self, n, norm, activation, num_domains)
        self.out1 = nn.Conv2d(2 * n, n_classes, 3, padding=1)
        self._pk_out1 = E.PackedConv()
        _init_weights(self, activation)

    def program(self, ctx, a, domain_label=None):
        x5, cats = self._encode(ctx, a, domain_label)
        return (E.head_conv(ctx, self._decode(ctx, x5, cats, domain_label), self.out1, self._pk_out1),)

    def forward(self, x, domain_label=None):
        return run_program(self, lambda ctx, a: self.program(ctx, a, domain_label), x)[0]


class Rec_Decoder(nn.Module):
    def __init__(self, n=16, num_classes=2, norm='bn', activation='relu', num_domains=None):
        super(Rec_Decoder, self).__init__()
        self.convu4 = ConvU_Rec(16 * n, norm, activation=activation, num_domains=num_domains)
        self.convu3 = ConvU_Rec(8 * n, norm, activation=activation, num_domains=num_domains)
        self.convu2 = ConvU_Rec(4 * n, norm, activation=activation, num_domains=num_domains)
        self.convu1 = ConvU_Rec(2 * n, norm, activation=activation, num_domains=num_domains)
        self.out1 = nn.Conv2d(n, num_classes, 3, padding=1)
        self._pk_out1 = E.PackedConv()
        _init_weights(self, activation)

    def program(self, ctx, a, domain_label=None):
        y = a
        for blk in (self.convu4, self.convu3, self.convu2, self.convu1):
            y = blk.run(ctx, y, domain_label)
        return (E.head_conv(ctx, y, self.out1, self._pk_out1),)

    def forward(self, x, domain_label=None):
        return run_program(self, lambda ctx, a: self.program(ctx, a, domain_label), x)[0]


class Unet2D_DS(_Head):
    def __init__(self, c=3, n=16, norm='bn', num_classes=2, activation='relu'):
        super(Unet2D_DS, self).__init__()
        _build_encoder(self, c, n, norm, activation)
        _build_decoder(self, n, norm, activation)
        self.seg5 = nn.Conv2d(16 * n, num_classes, 3, padding=1)
        self.seg4 = nn.Conv2d(16 * n, num_classes, 3, padding=1)
        self.seg3 = nn.Conv2d(8 * n, num_classes, 3, padding=1)
        self.seg2 = nn.Conv2d(4 * n, num_classes, 3, padding=1)
        self.seg1 = nn.Conv2d(2 * n, num_classes, 3, padding=1)
        self.upscore5 = nn.Upsample(scale_factor=16, mode='bilinear', align_corners=False)
        self.upscore4 = nn.Upsample(scale_factor=8, mode='bilinear', align_corners=False)
        self.upscore3 = nn.Upsample(scale_factor=4, mode='bilinear', align_corners=False)
        self.upscore2 = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
        self._pk_seg1 = E.PackedConv()
        _init_weights(self, activation)

    def forward(self, x, deep_sup=False):
        if deep_sup:
            raise NotImplementedError("deep supervision heads are outside the SSL-step hot path (SURVEY section 8); no sm_100a program")
        def program(ctx, a):
            x5, cats = self._encode(ctx, a)
            return (E.head_conv(ctx, self._decode(ctx, x5, cats), self.seg1, self._pk_seg1),)
        return run_program(self, program, x)[0]


class Unet2D_MS(_Head):
    def __init__(self, c=3, n=16, norm='bn', num_classes=2, activation='relu'):
        super(Unet2D_MS, self).__init__()
        _build_encoder(self, c, n, norm, activation)
        _build_decoder(self, n, norm, activation)
        self.seg5 = nn.Conv2d(16 * n, num_classes, 3, padding=1)
        self.seg4 = nn.Conv2d(16 * n, num_classes, 3, padding=1)
        self.seg3 = nn.Conv2d(8 * n, num_classes, 3, padding=1)
        self.seg2 = nn.Conv2d(4 * n, num_classes, 3, padding=1)
        self.seg1 = nn.Conv2d(2 * n, num_classes, 3, padding=1)
        self._pk_seg1 = E.PackedConv()
        _init_weights(self, activation)

    def forward(self, x, multi_scale_output=False):
        if multi_scale_output:
            raise NotImplementedError("multi-scale heads are outside the SSL-step hot path (SURVEY section 8); no sm_100a program")
        def program(ctx, a):
            x5, cats = self._encode(ctx, a)
            return (E.head_conv(ctx, self._decode(ctx, x5, cats), self.seg1, self._pk_seg1),)
        return run_program(self, program, x)[0]
