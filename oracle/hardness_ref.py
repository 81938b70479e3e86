"""TEST INFRASTRUCTURE (oracle): the reference's hardness bookkeeping of the unlabelled batch.

Restates /root/reference/utils/metrics.py ``dice_coefficient_numpy`` (:114-146), ``dice_coeff`` (:149-174),
``dice_coeff_2label`` (:176-201), ``dice_coeff_3label`` (:203-231) with ``ret_arr=True`` and the call site
train.py:705-718 (train_mnms.py uses the 3-label variant).  Pinned bit-for-bit against the reference's own
functions by ``oracle/make_golden.py`` (tests/golden/hardness.npz)."""
import numpy as np


def dice_coefficient(seg, gt):                                        # metrics.py:114-146
    seg = np.asarray(seg, dtype=bool)
    gt = np.asarray(gt, dtype=bool)
    inter = float(np.sum(np.logical_and(seg, gt).flatten()))
    s, g = float(np.sum(seg.flatten())), float(np.sum(gt.flatten()))
    if s == 0 and g == 0:
        return 0.0
    return (2 * inter + 1.0) / (1.001 + s + g)


def dice_parts(pred, target, mode):
    """List of per-sample Dice arrays, one per label part (the ``ret_arr=True`` return of the reference)."""
    pred, target = np.asarray(pred), np.asarray(target)
    if mode == "binary":                                               # metrics.py:149-174
        return [np.array([dice_coefficient(pred[i], target[i]) for i in range(pred.shape[0])])]
    if mode == "2label":                                               # metrics.py:176-201
        return [np.array([dice_coefficient(pred[i, c], target[i, c]) for i in range(pred.shape[0])]) for c in (0, 1)]
    if mode == "3label":                                               # metrics.py:203-231
        return [np.array([dice_coefficient((pred[i] == c).astype(float), (target[i] == c).astype(float)) for i in range(pred.shape[0])])
                for c in (1, 2, 3)]
    raise ValueError(mode)


def hardness(stu_pseudo_label, pseudo_label, mode="binary", first_epoch=False):   # train.py:705-718
    parts = dice_parts(stu_pseudo_label, pseudo_label, mode)
    tmp = parts[0].copy()
    for i in range(1, len(parts)):
        tmp += parts[i]
    h = 1 - tmp / len(parts)
    if first_epoch:
        for i in range(len(h)):
            h[i] = 1
    lq_idx, max_v = 0, -1
    for i in range(len(h)):
        if h[i] > max_v:
            max_v, lq_idx = h[i], i
    return h, lq_idx, np.stack(parts)
