"""Drop-in for the hot-path loss of the reference's utils/losses.py: ``DiceLossWithMask``
(losses.py:194-268; call sites train.py:521,816-836, train_mnms.py:479,732-754).

Same constructor and ``forward(inputs, target, mask=None, weight=None, softmax=False, sigmoid=False,
multi=False) -> 0-dim tensor`` contract, same quirks (class-0 Dice ignores the mask, smooth=1e-10,
mean over classes, ``multi`` = one global ratio), but the softmax / one-hot / masked sums / ratio run
as two fused sm_100a kernels with a device-side finalize (no per-class ``.item()`` host sync) and a
closed-form backward.  ``MaskedCEDice`` is the fully fused ``(ce*mask).mean() + dice`` term used by
the fused step (Tier B).  The file's other helpers upstream (FocalLoss, KL/MSE utilities, ...) are
not used by the training scripts and are not part of the hot path.
"""
import torch
import torch.nn as nn

from ustrun.loss_ops import FusedTerm, as_u8


def _prep(inputs, target, mask, softmax, sigmoid, multi, n_classes):
    if not inputs.is_cuda:
        raise RuntimeError("DiceLossWithMask runs on CUDA tensors only (sm_100a kernels, no CPU path)")
    if softmax and not multi:
        if target.dim() != 4 or target.size(1) != 1:
            raise AssertionError('predict & target shape do not match')
        tgt = target[:, 0]
        assert inputs.size(0) == tgt.size(0) and inputs.size()[2:] == tgt.size()[1:] and inputs.size(1) == n_classes, \
            'predict & target shape do not match'
        m = None
        if mask is not None:
            # losses.py:207-213: class c>=1 keeps a pixel iff mask*c == c  <=>  mask == 1
            m = as_u8(mask[:, 0] == 1).contiguous()
        return "softmax", as_u8(tgt).contiguous(), m
    if sigmoid and multi:
        tgt = target.squeeze(1)
        assert inputs.size() == tgt.size(), 'predict & target shape do not match'
        m = None
        if mask is not None:
            if not bool(((mask == 0) | (mask == 1)).all()):
                raise NotImplementedError("sigmoid/multi DiceLossWithMask kernel supports {0,1} masks (what the SSL step produces)")
            m = as_u8(mask.expand_as(tgt)).contiguous()
        return "sigmoid", as_u8(tgt).contiguous(), m
    raise NotImplementedError("DiceLossWithMask: only (softmax=True, multi=False) and (sigmoid=True, multi=True) -- the two "
                              "combinations train.py/train_mnms.py use -- have sm_100a kernels; there is no PyTorch fallback")


class DiceLossWithMask(nn.Module):
    def __init__(self, n_classes):
        super(DiceLossWithMask, self).__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, mask=None, weight=None, softmax=False, sigmoid=False, multi=False):
        if sigmoid and softmax:
            assert (0)
        branch, tgt, m = _prep(inputs, target, mask, softmax, sigmoid, multi, self.n_classes)
        cw = None
        if weight is not None and branch == "softmax":
            cw = torch.as_tensor(weight, dtype=torch.float32, device=inputs.device).contiguous()
        return FusedTerm.apply(inputs, tgt, m, branch, 0.0, 1.0, cw)


class MaskedCEDice(nn.Module):
    """``(ce(logits, target) * mask).mean() + DiceLossWithMask(...)`` in one fused node
    (train.py:816-817,829-836).  softmax branch: target [B,H,W], mask [B,1,H,W] or None;
    sigmoid branch: target/mask [B,C,H,W]."""

    def __init__(self, n_classes, branch="softmax"):
        super().__init__()
        self.n_classes, self.branch = n_classes, branch

    def forward(self, logits, target, mask=None):
        if self.branch == "softmax":
            m = None if mask is None else as_u8(mask.reshape(target.shape)).contiguous()
        else:
            m = None if mask is None else as_u8(mask.expand_as(target)).contiguous()
        return FusedTerm.apply(logits, as_u8(target).contiguous(), m, self.branch, 1.0, 1.0, None)
