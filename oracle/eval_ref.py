"""TEST INFRASTRUCTURE (oracle): evaluation-side restatement of the reference.

Label encodings: train.py:590-608 (fundus / prostate / BUSI; identical lines in ``test()``, train.py:281-288) and
train_mnms.py:549-556.  Prediction rule: train.py:295-302.  Dice: utils/metrics.py (see oracle/hardness_ref.py).
dc / jc: ``medpy.metric.binary.dc`` / ``.jc`` (train.py:307-311).  medpy is a third-party dependency of the reference
that is neither vendored under /root/reference nor installed here (the reference pins no version; current release
0.5.2), so its published definitions are restated: dc = 2|A & B| / (|A| + |B|) with 0.0 on ZeroDivisionError,
jc = |A & B| / |A | B| (medpy lets the ZeroDivisionError of an empty union propagate; 0.0 here).

Pinning: label encodings, the prediction rule and the per-sample Dice values are checked against the reference's own
lines / functions (oracle/make_golden.py ``case_eval`` -> tests/golden/eval.npz, re-checked by tests/test_oracle_golden.py).
The batch-mean Dice agrees to the last bit or one ulp: the reference sums Python floats with ``sum()`` (Neumaier-compensated
since Python 3.12), this file and the device kernel add sequentially in double.  dc / jc: PARITY UNPINNED (no medpy here)."""
import numpy as np
import torch

from . import hardness_ref as Hr


def encode_labels(y: torch.Tensor, dataset: str) -> torch.Tensor:
    if dataset == "fundus":                                            # train.py:591-593
        cup, disc = y.eq(0).float(), y.le(128).float()
        return torch.cat((cup.unsqueeze(1), disc.unsqueeze(1)), dim=1)
    if dataset == "prostate":                                          # train.py:600
        return y.eq(0).long()
    if dataset == "BUSI":                                              # train.py:605
        return y.eq(255).long()
    if dataset == "mnms":                                              # train_mnms.py:549-552
        m = y[:, ..., 0].eq(255).float()
        m[y[:, ..., 1].eq(255)] = 2
        m[y[:, ..., 2].eq(255)] = 3
        return m.long()
    raise ValueError(dataset)


def predict(output: torch.Tensor, dataset: str) -> torch.Tensor:      # train.py:295-302
    if dataset == "fundus":
        return torch.sigmoid(output).ge(0.5)
    return torch.max(torch.softmax(output, dim=1), dim=1)[1]


def binary_dc(a, b):                                                   # medpy.metric.binary.dc
    a, b = np.asarray(a, dtype=bool), np.asarray(b, dtype=bool)
    # int(): numpy >= 2 returns a numpy scalar here, which turns medpy's ZeroDivisionError branch into a silent nan
    inter, s1, s2 = int(np.count_nonzero(a & b)), int(np.count_nonzero(a)), int(np.count_nonzero(b))
    try:
        return 2.0 * inter / float(s1 + s2)
    except ZeroDivisionError:
        return 0.0


def binary_jc(a, b):                                                   # medpy.metric.binary.jc (0.0 instead of raising)
    a, b = np.asarray(a, dtype=bool), np.asarray(b, dtype=bool)
    union = int(np.count_nonzero(a | b))
    return float(int(np.count_nonzero(a & b))) / float(union) if union else 0.0


def seg_metrics(pred_label, mask, dataset):
    """dice: train.py:303 (``dice_calcu[dataset](pred, mask)`` = batch means); dc / jc: train.py:305-320."""
    mode = {"prostate": "binary", "BUSI": "binary", "fundus": "2label", "mnms": "3label"}[dataset]
    pred_label, mask = np.asarray(pred_label), np.asarray(mask)
    parts = Hr.dice_parts(pred_label, mask, mode)
    dice = [sum(p) / len(p) for p in parts]
    n = pred_label.shape[0]
    if mode == "binary":
        po, mo = pred_label[:, None], mask[:, None]
    elif mode == "2label":
        po, mo = pred_label, mask
    else:
        po = np.stack([pred_label == c for c in (1, 2, 3)], 1)
        mo = np.stack([mask == c for c in (1, 2, 3)], 1)
    dc, jc = [0.0] * po.shape[1], [0.0] * po.shape[1]
    for j in range(n):
        for i in range(po.shape[1]):
            dc[i] += binary_dc(po[j, i], mo[j, i])
            jc[i] += binary_jc(po[j, i], mo[j, i])
    return np.array(dice), np.array([d / n for d in dc]), np.array([v / n for v in jc])
