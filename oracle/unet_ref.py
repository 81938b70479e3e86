"""Oracle (test infrastructure): functional torch-CPU restatement of the reference networks.

Every function works on a flat ``state`` dict keyed exactly like the reference ``state_dict()``
(so weights can be exchanged with the reference modules and with the CUDA drop-ins) and uses only
``torch.nn.functional`` primitives.  Nothing here is imported by the product path.

Reference lines followed:
  * UNet-A  ``networks/unet_model.py:6-39`` + ``networks/unet_parts.py:8-77``
  * UNet-B  ``networks/unet.py:32-117`` (ConvD/ConvU), ``:298-334`` (UNet), ``:248-296`` (Encoder/Decoder)
  * DSBN    ``networks/dsbn.py:4-34``; Rec decoder ``networks/unet.py:120-165,339-362``
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------
# state construction (same layer-constructor order as the reference => same RNG stream)
# --------------------------------------------------------------------------------------------
def _put(state, prefix, mod):
    for k, v in mod.state_dict().items():
        state[f"{prefix}.{k}"] = v.detach().clone()


def _double_conv_state(state, prefix, cin, cout, cmid=None, norm="bn", num_domains=None):
    # unet_parts.py:11-22 : conv(no bias) bn relu conv(no bias) bn relu -> indices 0,1,3,4
    cmid = cmid or cout
    _put(state, f"{prefix}.0", nn.Conv2d(cin, cmid, 3, padding=1, bias=False))
    _norm_state(state, f"{prefix}.1", cmid, norm, num_domains)
    _put(state, f"{prefix}.3", nn.Conv2d(cmid, cout, 3, padding=1, bias=False))
    _norm_state(state, f"{prefix}.4", cout, norm, num_domains)


def init_unet_a(n_channels, n_classes, bilinear=False, seed=None, norm="bn", num_domains=None):
    """State of ``unet_model.UNet(n_channels, n_classes, bilinear)`` (unet_model.py:7-23).
    ``norm='dsbn'`` is NOT upstream: UNet-A with every BatchNorm2d replaced by the reference's
    ``DomainSpecificBatchNorm2d`` (dsbn.py:4-34; keys ``....1.bns.{d}.*``) -- the tensor-bound model of
    BASELINE.json configs[2] ("DSBN over 3 domains").  BatchNorm construction draws no random numbers, so
    the conv weights equal those of the plain model for the same seed."""
    if seed is not None:
        torch.manual_seed(seed)
    st = OrderedDict()
    kw = dict(norm=norm, num_domains=num_domains)
    _double_conv_state(st, "inc.double_conv", n_channels, 64, **kw)
    factor = 2 if bilinear else 1
    widths = [(64, 128), (128, 256), (256, 512), (512, 1024 // factor)]
    for i, (a, b) in enumerate(widths, 1):
        _double_conv_state(st, f"down{i}.maxpool_conv.1.double_conv", a, b, **kw)
    ups = [(1024, 512 // factor), (512, 256 // factor), (256, 128 // factor), (128, 64)]
    for i, (a, b) in enumerate(ups, 1):
        if bilinear:
            _double_conv_state(st, f"up{i}.conv.double_conv", a, b, a // 2, **kw)
        else:
            _put(st, f"up{i}.up", nn.ConvTranspose2d(a, a // 2, kernel_size=2, stride=2))
            _double_conv_state(st, f"up{i}.conv.double_conv", a, b, **kw)
    _put(st, "outc.conv", nn.Conv2d(64, n_classes, kernel_size=1))
    return st


def _norm_state(state, prefix, planes, norm, num_domains=None):
    if norm == "bn":
        _put(state, prefix, nn.BatchNorm2d(planes))
    elif norm == "dsbn":
        for d in range(num_domains):
            _put(state, f"{prefix}.bns.{d}", nn.BatchNorm2d(planes))
    else:
        raise ValueError("Normalization type {} is not supporter".format(norm))


def _kaiming_fan_out(state, activation="relu"):
    # unet.py:314-319 : kaiming_normal_(fan_out) on every conv weight; BN weight=1, bias=0
    for k, v in state.items():
        if k.endswith(".weight") and v.dim() == 4:
            nn.init.kaiming_normal_(v, mode="fan_out", nonlinearity=activation)
        elif k.endswith(".weight") and v.dim() == 1:
            v.fill_(1.0)
        elif k.endswith(".bias") and k.rsplit(".", 1)[0] + ".running_mean" in state:
            v.fill_(0.0)


def _convd_state(st, p, cin, c, norm, num_domains=None):
    for j, ci in ((1, cin), (2, c), (3, c)):
        _put(st, f"{p}.conv{j}", nn.Conv2d(ci, c, 3, 1, 1, bias=True))
        _norm_state(st, f"{p}.bn{j}", c, norm, num_domains)


def _convu_state(st, p, planes, norm, first, num_domains=None):
    if not first:
        _put(st, f"{p}.conv1", nn.Conv2d(2 * planes, planes, 3, 1, 1, bias=True))
        _norm_state(st, f"{p}.bn1", planes, norm, num_domains)
    _put(st, f"{p}.conv2", nn.Conv2d(planes, planes // 2, 1, 1, 0, bias=True))
    _norm_state(st, f"{p}.bn2", planes // 2, norm, num_domains)
    _put(st, f"{p}.conv3", nn.Conv2d(planes, planes, 3, 1, 1, bias=True))
    _norm_state(st, f"{p}.bn3", planes, norm, num_domains)


def init_unet_b(n_channels=3, n_classes=2, n=16, norm="bn", seed=None, head="out1",
                encoder=True, decoder=True, num_domains=None):
    """State of ``networks.unet.UNet`` (unet.py:299-319); ``Encoder``/``Decoder`` via flags.
    ``norm='dsbn', num_domains=k``: the patched network of SURVEY A2(ii) (upstream raises TypeError at
    construction because ConvD/ConvU do not pass ``num_domains`` on, unet.py:38,82)."""
    if seed is not None:
        torch.manual_seed(seed)
    st = OrderedDict()
    nd = num_domains
    if encoder:
        _convd_state(st, "convd1", n_channels, n, norm, nd)
        _convd_state(st, "convd2", n, 2 * n, norm, nd)
        _convd_state(st, "convd3", 2 * n, 4 * n, norm, nd)
        _convd_state(st, "convd4", 4 * n, 8 * n, norm, nd)
        _convd_state(st, "convd5", 8 * n, 16 * n, norm, nd)
    if decoder:
        _convu_state(st, "convu4", 16 * n, norm, True, nd)
        _convu_state(st, "convu3", 8 * n, norm, False, nd)
        _convu_state(st, "convu2", 4 * n, norm, False, nd)
        _convu_state(st, "convu1", 2 * n, norm, False, nd)
        _put(st, head, nn.Conv2d(2 * n, n_classes, 3, padding=1))
    _kaiming_fan_out(st)
    return st


def init_rec_decoder(n=16, num_classes=2, norm="bn", num_domains=None, seed=None):
    """State of ``networks.unet.Rec_Decoder`` (unet.py:340-354, ConvU_Rec :121-132)."""
    if seed is not None:
        torch.manual_seed(seed)
    st = OrderedDict()
    for name, planes in (("convu4", 16 * n), ("convu3", 8 * n), ("convu2", 4 * n), ("convu1", 2 * n)):
        h = planes // 2
        _put(st, f"{name}.conv1", nn.Conv2d(planes, h, 3, 1, 1, bias=True))
        _norm_state(st, f"{name}.bn1", h, norm, num_domains)
        _put(st, f"{name}.conv2", nn.Conv2d(h, h, 1, 1, 0, bias=True))
        _norm_state(st, f"{name}.bn2", h, norm, num_domains)
        _put(st, f"{name}.conv3", nn.Conv2d(h, h, 3, 1, 1, bias=True))
        _norm_state(st, f"{name}.bn3", h, norm, num_domains)
    _put(st, "out1", nn.Conv2d(n, num_classes, 3, padding=1))
    _kaiming_fan_out(st)
    return st


def split_state(state):
    """(params, buffers): params are what ``parameters()`` yields, in order."""
    params, buffers = OrderedDict(), OrderedDict()
    for k, v in state.items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            buffers[k] = v
        else:
            params[k] = v
    return params, buffers


# --------------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------------
def _bn(st, p, x, training):
    """Train-mode BatchNorm2d with running-stat side effects (torch semantics: biased var for
    normalisation, unbiased for running_var, momentum 0.1, eps 1e-5, num_batches_tracked += 1)."""
    rm, rv = st[p + ".running_mean"], st[p + ".running_var"]
    if training:
        st[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, st[p + ".weight"], st[p + ".bias"], training, BN_MOMENTUM, BN_EPS)


def _dsbn(st, p, x, domain_label, training):
    # dsbn.py:24-27 : the WHOLE batch goes through bns[domain_label[0]]
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    return _bn(st, f"{p}.bns.{int(domain_label[0])}", x, training)


def _double_conv(st, p, x, training, domain_label=None):
    x = F.relu(_norm(st, p + ".1", F.conv2d(x, st[p + ".0.weight"], None, 1, 1), training, domain_label))
    x = F.relu(_norm(st, p + ".4", F.conv2d(x, st[p + ".3.weight"], None, 1, 1), training, domain_label))
    return x


def unet_a_forward(st, x, training=True, feature=False, bilinear=False, domain_label=None):
    """``unet_model.UNet.forward`` (unet_model.py:25-39).  ``domain_label`` only for the DSBN extension
    (``init_unet_a(norm='dsbn')``); None = upstream."""
    x1 = _double_conv(st, "inc.double_conv", x, training, domain_label)
    skips = [x1]
    h = x1
    for i in range(1, 5):
        h = _double_conv(st, f"down{i}.maxpool_conv.1.double_conv", F.max_pool2d(h, 2), training, domain_label)
        skips.append(h)
    for i in range(1, 5):
        skip = skips[4 - i]
        if bilinear:
            h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        else:
            h = F.conv_transpose2d(h, st[f"up{i}.up.weight"], st[f"up{i}.up.bias"], stride=2)
        dy, dx = skip.size(2) - h.size(2), skip.size(3) - h.size(3)
        h = F.pad(h, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])      # unet_parts.py:62-63
        h = _double_conv(st, f"up{i}.conv.double_conv", torch.cat([skip, h], 1), training, domain_label)
    logits = F.conv2d(h, st["outc.conv.weight"], st["outc.conv.bias"])
    return (logits, h) if feature else logits


def _act(x, activation):
    return F.relu(x) if activation == "relu" else F.leaky_relu(x, 0.01)


def _norm(st, p, x, training, domain_label=None):
    if p + ".weight" in st:
        return _bn(st, p, x, training)
    if domain_label is None:
        raise TypeError("forward() missing 1 required positional argument: 'domain_label'")
    return _dsbn(st, p, x, domain_label, training)


def _convd(st, p, x, first, training, activation="relu", domain_label=None):
    # unet.py:52-72 : note NO activation after bn1
    if not first:
        x = F.max_pool2d(x, 2)
    x = _norm(st, p + ".bn1", F.conv2d(x, st[p + ".conv1.weight"], st[p + ".conv1.bias"], 1, 1), training, domain_label)
    y = _act(_norm(st, p + ".bn2", F.conv2d(x, st[p + ".conv2.weight"], st[p + ".conv2.bias"], 1, 1), training, domain_label), activation)
    z = _act(_norm(st, p + ".bn3", F.conv2d(y, st[p + ".conv3.weight"], st[p + ".conv3.bias"], 1, 1), training, domain_label), activation)
    return z


def _convu(st, p, x, prev, first, training, activation="relu", domain_label=None):
    # unet.py:96-117
    if not first:
        x = _act(_norm(st, p + ".bn1", F.conv2d(x, st[p + ".conv1.weight"], st[p + ".conv1.bias"], 1, 1), training, domain_label), activation)
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    y = _act(_norm(st, p + ".bn2", F.conv2d(y, st[p + ".conv2.weight"], st[p + ".conv2.bias"]), training, domain_label), activation)
    y = torch.cat([prev, y], 1)
    y = _act(_norm(st, p + ".bn3", F.conv2d(y, st[p + ".conv3.weight"], st[p + ".conv3.bias"], 1, 1), training, domain_label), activation)
    return y


def unet_b_encoder(st, x, training=True, activation="relu", domain_label=None):
    feats = []
    h = x
    for i in range(1, 6):
        h = _convd(st, f"convd{i}", h, i == 1, training, activation, domain_label)
        feats.append(h)
    return feats


def unet_b_decoder(st, feats, training=True, activation="relu", head="out1", domain_label=None):
    y = _convu(st, "convu4", feats[-1], feats[-2], True, training, activation, domain_label)
    y = _convu(st, "convu3", y, feats[-3], False, training, activation, domain_label)
    y = _convu(st, "convu2", y, feats[-4], False, training, activation, domain_label)
    y = _convu(st, "convu1", y, feats[-5], False, training, activation, domain_label)
    return F.conv2d(y, st[head + ".weight"], st[head + ".bias"], 1, 1)


def unet_b_forward(st, x, training=True, activation="relu", head="out1", domain_label=None):
    """``networks.unet.UNet.forward`` (unet.py:321-334).  ``domain_label`` is the documented
    6-line patch of SURVEY A2(ii) (threading DSBN through ConvD/ConvU); None = upstream."""
    return unet_b_decoder(st, unet_b_encoder(st, x, training, activation, domain_label), training,
                          activation, head, domain_label)


def rec_decoder_forward(st, x, domain_label=None, training=True, activation="relu"):
    """``Rec_Decoder.forward`` (unet.py:356-362) with ``ConvU_Rec.forward`` (unet.py:139-165)."""
    h = x
    for name in ("convu4", "convu3", "convu2", "convu1"):
        def nrm(j, t):
            p = f"{name}.bn{j}"
            if p + ".weight" in st:                       # plain bn
                return _bn(st, p, t, training)
            if domain_label is None:                      # unet.py:142-145 -> TypeError upstream
                raise TypeError("forward() missing 1 required positional argument: 'domain_label'")
            return _dsbn(st, p, t, domain_label, training)
        h = _act(nrm(1, F.conv2d(h, st[f"{name}.conv1.weight"], st[f"{name}.conv1.bias"], 1, 1)), activation)
        y = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)
        y = _act(nrm(2, F.conv2d(y, st[f"{name}.conv2.weight"], st[f"{name}.conv2.bias"])), activation)
        h = _act(nrm(3, F.conv2d(y, st[f"{name}.conv3.weight"], st[f"{name}.conv3.bias"], 1, 1)), activation)
    return F.conv2d(h, st["out1.weight"], st["out1.bias"], 1, 1)


def conv_flops_unet_a(n_channels, n_classes, H, W):
    """Algorithmic conv+convT FLOPs (MAC x 2) per image for UNet-A."""
    f = 0
    def dc(cin, cout, h, w):
        return 2 * 9 * h * w * (cin * cout + cout * cout)
    f += dc(n_channels, 64, H, W)
    c, h, w = 64, H, W
    for _ in range(4):
        h, w = h // 2, w // 2
        f += dc(c, 2 * c, h, w)
        c *= 2
    for _ in range(4):
        f += 2 * h * w * c * (c // 2) * 4          # convT k2 s2
        h, w = h * 2, w * 2
        f += dc(c, c // 2, h, w)
        c //= 2
    f += 2 * H * W * 64 * n_classes
    return f
