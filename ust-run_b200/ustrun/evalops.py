"""Evaluation-side pieces of the reference on the device: the per-dataset label encodings
(train.py:590-608, :281-288; train_mnms.py:549-556), the prediction rule of ``test()`` (train.py:295-302) and the
per-part Dice / dc / jc batch means (train.py:303-320) -- everything of ``test()`` except medpy's hd95 / asd, which
need distance transforms and stay on the host.  ``evaluate_batch`` strings them together around an eval-mode forward
(running-statistics BatchNorm) and the fused CE + Dice loss; nothing is copied to the host."""
from __future__ import annotations

import torch

from . import _lib as L
from .engine import _call, _ptr, _stream
from .loss_ops import as_u8, term_forward

LABEL_MODES = {"prostate": 0, "BUSI": 1, "fundus": 2, "mnms": 3}
METRIC_MODES = {"prostate": 0, "BUSI": 0, "fundus": 1, "mnms": 2}


def encode_labels(y: torch.Tensor, dataset: str) -> torch.Tensor:
    """float label image [B,H,W] ([B,H,W,3] for "mnms") -> uint8 targets ([B,2,H,W] for "fundus")."""
    L.require_device()
    mode = LABEL_MODES[dataset]
    y = y.float().contiguous()
    if mode == 3:
        if y.dim() != 4 or y.shape[-1] != 3:
            raise ValueError("encode_labels('mnms'): expected [B,H,W,3]")
        B, H, W = y.shape[0], y.shape[1], y.shape[2]
    else:
        if y.dim() != 3:
            raise ValueError("encode_labels: expected [B,H,W]")
        B, H, W = y.shape
    out = torch.empty((B, 2, H, W) if mode == 2 else (B, H, W), dtype=torch.uint8, device=y.device)
    _call("ustrun_encode_labels", _ptr(y), mode, B, H, W, _ptr(out), _stream())
    return out


def predict(logits: torch.Tensor, branch: str) -> torch.Tensor:
    """train.py:295-302: "softmax" -> argmax of the softmax probabilities [B,H,W]; "sigmoid" -> sigmoid >= 0.5 [B,C,H,W]."""
    L.require_device()
    logits = logits.float().contiguous()
    B, C, H, W = logits.shape
    sig = branch == "sigmoid"
    pred = torch.empty((B, C, H, W) if sig else (B, H, W), dtype=torch.uint8, device=logits.device)
    _call("ustrun_predict", _ptr(logits), 1 if sig else 0, B, C, H, W, _ptr(pred), _stream())
    return pred


def seg_metrics(pred: torch.Tensor, target: torch.Tensor, dataset: str):
    """Batch means per label part: dict(dice=, dc=, jc=) of float64 device tensors [parts]."""
    L.require_device()
    mode = METRIC_MODES[dataset]
    p, t = as_u8(pred).contiguous(), as_u8(target).contiguous()
    if p.shape != t.shape:
        raise ValueError("seg_metrics: prediction and target shapes differ")
    B, H, W = p.shape[0], p.shape[-2], p.shape[-1]
    parts = (1, 2, 3)[mode]
    ws = torch.empty(B * 9, dtype=torch.int32, device=p.device)
    out = torch.empty((3, parts), dtype=torch.float64, device=p.device)
    _call("ustrun_seg_metrics", _ptr(p), _ptr(t), B, H, W, mode, _ptr(ws), _ptr(out), _stream())
    return {"dice": out[0], "dc": out[1], "jc": out[2]}


@torch.no_grad()
def evaluate_batch(model, data: torch.Tensor, label: torch.Tensor, dataset: str):
    """One batch of ``test()`` (train.py:276-320) on the device: label encoding, eval-mode forward, ``loss_seg`` =
    CE.mean() + DiceLossWithMask (BCE / multi-Dice for fundus), prediction, Dice / dc / jc batch means."""
    branch = "sigmoid" if dataset == "fundus" else "softmax"
    target = encode_labels(label, dataset)
    was_training = model.training
    model.eval()
    try:
        logits = model(data)
    finally:
        model.train(was_training)
    loss3, _ = term_forward(logits.float().contiguous(), target, None, branch)
    pred = predict(logits, branch)
    m = seg_metrics(pred, target, dataset)
    m.update(loss_seg=loss3[0], pred=pred, target=target, logits=logits)
    return m
