"""Fused CE+Dice loss ops on libustrun_sm100.so (K13/K14 of SURVEY 2.2).

``term_forward`` runs pass 1 (one read of the logits: softmax once per pixel, 3C+1 partial sums,
device-side finalize -> loss scalars + pass-2 coefficients, NO host sync, unlike the reference's
per-class ``.item()`` at utils/losses.py:261); ``term_backward`` runs pass 2 (closed-form gradient,
SURVEY App. C) writing dLoss/dlogits directly.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from .engine import _call, _profiled, _ptr, _stream


def as_u8(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.uint8 else t.to(torch.uint8)


def term_forward(logits, target_u8, mask_u8, branch, ce_w=1.0, dice_w=1.0, class_weight=None, allreduce=None, world=1):
    """logits fp32 NCHW [B,C,H,W]; softmax: target/mask uint8 [B,H,W]; sigmoid: uint8 [B,C,H,W].
    Returns (loss3, coef): loss3 = [ce_w*CE + dice_w*Dice, CE, Dice] (device tensor).
    ``allreduce`` (data parallel): callable summing a tensor across ranks in place; the (3C+1) partial
    sums are reduced across ranks before the finalize kernel, so loss and gradient coefficients are
    those of the GLOBAL batch (world x local pixels)."""
    L.require_device()
    B, C, H, W = logits.shape
    dev = logits.device
    loss3 = torch.empty(3, dtype=torch.float32, device=dev)
    if allreduce is not None:
        ncols = 3 * C + 1 if branch == "softmax" else 4
        ws = torch.empty(L.MAX_PARTS * ncols, dtype=torch.float32, device=dev)
        sums = torch.empty(ncols, dtype=torch.float32, device=dev)
        nparts = ctypes.c_int(0)
        if branch == "softmax":
            coef = torch.empty(4 * C + 4, dtype=torch.float32, device=dev)
            _call("ustrun_ce_dice_softmax_partials", _ptr(logits), _ptr(target_u8), _ptr(mask_u8), B, C, H, W, _ptr(ws), ctypes.byref(nparts), _stream())
            _call("ustrun_reduce_rows", _ptr(ws), nparts.value, ncols, _ptr(sums), _stream())
            allreduce(sums)
            _call("ustrun_ce_dice_softmax_finalize", _ptr(sums), 1, C, float(B * H * W) * world, float(ce_w), float(dice_w), _ptr(class_weight),
                  _ptr(coef), _ptr(loss3), _stream())
        else:
            coef = torch.empty(4, dtype=torch.float32, device=dev)
            _call("ustrun_bce_dice_sigmoid_partials", _ptr(logits), _ptr(target_u8), _ptr(mask_u8), B, C, H, W, _ptr(ws), ctypes.byref(nparts), _stream())
            _call("ustrun_reduce_rows", _ptr(ws), nparts.value, ncols, _ptr(sums), _stream())
            allreduce(sums)
            _call("ustrun_bce_dice_sigmoid_finalize", _ptr(sums), 1, float(B * C * H * W) * world, float(ce_w), float(dice_w), _ptr(coef), _ptr(loss3), _stream())
        return loss3, coef
    # algorithmic bytes of pass 1: the logits once (fp32) + target / mask bytes
    nlab = B * H * W * (1 if branch == "softmax" else C)
    p1_bytes = B * C * H * W * 4 + nlab * (2 if mask_u8 is not None else 1)
    if branch == "softmax":
        ws = torch.empty(L.MAX_PARTS * (3 * C + 1), dtype=torch.float32, device=dev)
        coef = torch.empty(4 * C + 4, dtype=torch.float32, device=dev)
        _profiled("hbm_ce_dice_pass1", p1_bytes, "ustrun_ce_dice_softmax_fwd", _ptr(logits), _ptr(target_u8), _ptr(mask_u8), B, C, H, W, float(ce_w), float(dice_w),
              _ptr(class_weight), _ptr(ws), _ptr(coef), _ptr(loss3), _stream())
    else:
        ws = torch.empty(L.MAX_PARTS * 4, dtype=torch.float32, device=dev)
        coef = torch.empty(4, dtype=torch.float32, device=dev)
        _profiled("hbm_ce_dice_pass1", p1_bytes, "ustrun_bce_dice_sigmoid_fwd", _ptr(logits), _ptr(target_u8), _ptr(mask_u8), B, C, H, W, float(ce_w), float(dice_w),
              _ptr(ws), _ptr(coef), _ptr(loss3), _stream())
    return loss3, coef


def term_backward(logits, target_u8, mask_u8, branch, coef, upstream=None, gscale=1.0, out=None, accumulate=False):
    """dlogits (+)= gscale * upstream * dLoss/dlogits  (upstream: 0-dim device tensor or None)."""
    B, C, H, W = logits.shape
    if out is None:
        out = torch.empty_like(logits)
        accumulate = False
    name = "ustrun_ce_dice_softmax_bwd" if branch == "softmax" else "ustrun_bce_dice_sigmoid_bwd"
    nlab = B * H * W * (1 if branch == "softmax" else C)
    _profiled("hbm_ce_dice_pass2", 2 * B * C * H * W * 4 + nlab * (2 if mask_u8 is not None else 1), name, _ptr(logits), _ptr(target_u8), _ptr(mask_u8), B, C, H, W, _ptr(coef), _ptr(upstream), float(gscale), _ptr(out),
          1 if accumulate else 0, _stream())
    return out


class FusedTerm(torch.autograd.Function):
    """autograd node: loss = ce_w * mean(CE*mask) + dice_w * DiceLossWithMask (one loss term)."""

    @staticmethod
    def forward(ctx, logits, target_u8, mask_u8, branch, ce_w, dice_w, class_weight):
        logits = logits.detach()
        if logits.dtype != torch.float32 or not logits.is_contiguous():
            logits = logits.float().contiguous()
        loss3, coef = term_forward(logits, target_u8, mask_u8, branch, ce_w, dice_w, class_weight)
        ctx.save_for_backward(logits, target_u8, mask_u8 if mask_u8 is not None else torch.empty(0, device=logits.device), coef)
        ctx.branch, ctx.has_mask = branch, mask_u8 is not None
        return loss3[0]

    @staticmethod
    def backward(ctx, g):
        logits, target_u8, mask_u8, coef = ctx.saved_tensors
        up = g.detach().float().contiguous()
        d = term_backward(logits, target_u8, mask_u8 if ctx.has_mask else None, ctx.branch, coef, upstream=up)
        return d, None, None, None, None, None, None
