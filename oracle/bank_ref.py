"""Oracle (test infrastructure): restatement of the reference's confidence-bank bookkeeping, the inline code of
``train.py`` that follows the per-sample hardness (second half).

Reference lines followed:
  * bank update / adaptive threshold   ``train.py:754-781``  (``simple_ulb_idx = hardness < choice_th`` ... ``choice_th = min(args.increase*choice_th, 0.1)``)
  * CutMix partner pool + choice       ``train.py:612-625``
  * low-quality sample bookkeeping     ``train.py:741-743`` (``lq_u / lq_pl / lq_mask``), its CutMix with a labelled image
    inside the box that covers both structures ``train.py:722-739`` and ``obtain_all_cover_box`` ``train.py:242-251``

All host randomness enters as inputs (``choice`` draws, ``new_choice``).  State is a plain dict so that it can be compared with
the device-side ``ustrun.bank.ConfidenceBank``.  Pinned by oracle/make_golden.py::case_bank, which executes the reference's own
lines (sliced out of train.py, unmodified) on the same inputs."""
from __future__ import annotations

import numpy as np
import torch


def new_state(choice_th=0.1):
    return dict(simple_ulb=None, cor_pl=None, cor_mask=None, cor_hardness=np.zeros(0, np.float64), choice_th=float(choice_th))


def bank_update(st, hardness, ulb_x_w, pseudo_label, mask, max_len=10, increase=1.0005):
    """train.py:745-779.  hardness: float64 numpy [Bu]; tensors are the unlabelled batch's weak view, the teacher's pseudo
    label and confidence mask.  Returns the new state (the input dict is not modified)."""
    st = dict(st)
    choice_th = st["choice_th"]
    sel = hardness < choice_th                                                   # :745
    cur = int(sel.astype(int).sum())                                             # :746
    sel_t = torch.from_numpy(sel)
    if st["simple_ulb"] is None or len(st["simple_ulb"]) == 0:                   # :747
        st["simple_ulb"] = ulb_x_w[sel_t].clone()
        st["cor_pl"] = pseudo_label[sel_t].clone()
        st["cor_hardness"] = hardness[sel].copy()
        st["cor_mask"] = mask[sel_t].clone()
        if len(st["simple_ulb"]) > 0:                                            # :754
            choice_th = min(choice_th, st["cor_hardness"].max())
    else:
        if cur > 0:                                                              # :757
            n_old = len(st["simple_ulb"])
            newlen = max_len - cur if n_old + cur > max_len else n_old           # :758-761
            st["simple_ulb"] = torch.cat((ulb_x_w[sel_t].clone(), st["simple_ulb"][:newlen]), dim=0)
            st["cor_pl"] = torch.cat((pseudo_label[sel_t].clone(), st["cor_pl"][:newlen]), dim=0)
            st["cor_hardness"] = np.concatenate((hardness[sel].copy(), st["cor_hardness"][:newlen]))
            st["cor_mask"] = torch.cat((mask[sel_t].clone(), st["cor_mask"][:newlen]), dim=0)
            choice_th = min(choice_th, st["cor_hardness"].max())                 # :768
        else:
            choice_th = min(increase * choice_th, 0.1)                           # :770
    st["choice_th"] = float(choice_th)
    return st


def cut_pool(st, lb_x_w, lb_mask, lb_mask_shape):
    """train.py:612-620 -> (cut_img, cut_label, cut_mask)."""
    if st["simple_ulb"] is None or len(st["simple_ulb"]) == 0:
        return lb_x_w.clone(), lb_mask.clone(), torch.ones(lb_mask_shape)
    return (torch.cat((lb_x_w.clone(), st["simple_ulb"]), dim=0), torch.cat((lb_mask.clone(), st["cor_pl"]), dim=0),
            torch.cat((torch.ones(lb_mask_shape), st["cor_mask"]), dim=0))


def draw_choice(n_bank, B_l, B_u, r_lb, r_u, perm):
    """train.py:615,621-625 with the host's random draws as inputs: ``r_lb`` int [B_u] uniform in [0, B_l), ``r_u`` float64
    [B_u] uniform in [0, 1) (mapped to a bank slot), ``perm`` a permutation of range(B_u)."""
    if n_bank == 0:
        return np.asarray(r_lb[:B_u], dtype=np.int64)
    k = min(int(B_u * 0.5), n_bank)
    in_lb = np.asarray(r_lb[:B_u - k], dtype=np.int64)
    in_simple = B_l + np.floor(np.asarray(r_u[:k], dtype=np.float64) * n_bank).astype(np.int64)
    return np.concatenate((in_lb, in_simple))[np.asarray(perm)]


def all_cover_box(region):
    """train.py:242-251 for a non-empty region: rows from the first to the last non-zero pixel (row-major order), columns from
    the smallest to the largest non-zero column."""
    loc = torch.nonzero(region).numpy()
    if len(loc) == 0:
        raise ValueError("empty region: the reference falls back to a random CutMix box (host RNG)")
    box = torch.zeros_like(region)
    y1, y2, x1, x2 = loc[0, 0], loc[-1, 0], loc[:, 1].min(), loc[:, 1].max()
    box[y1:y2 + 1, x1:x2 + 1] = 1
    return box


def lq_region(lq_pl, lb_mask, new_choice, dataset):
    """train.py:722-730: union of the low-quality sample's pseudo label and the chosen labelled mask."""
    if dataset == "fundus":
        region = lq_pl[0, 1].clone()
        region[lq_pl[0, 0].long() == 1] = 1
        region[lb_mask[new_choice, 0].long() == 1] = 1
        region[lb_mask[new_choice, 1].long() == 1] = 1
    else:
        region = lq_pl[0].clone()
        region[lb_mask[new_choice].long() > 0] = 1
    return region


def lq_compose(lq_u, lq_pl, lq_mask, lb_x_w, lb_mask, new_choice, dataset):
    """train.py:720-738 -> (lq_s, pseudo_label_lq, mask_lq, box)."""
    box = all_cover_box(lq_region(lq_pl, lb_mask, new_choice, dataset)).unsqueeze(0)
    img_box = box.unsqueeze(1)
    label_box = box.unsqueeze(1) if dataset == "fundus" else box
    lq_s = lq_u * (1 - img_box) + lb_x_w[[new_choice]] * img_box
    pl = (lq_pl * (1 - label_box) + lb_mask[[new_choice]] * label_box).long()
    if dataset in ("fundus", "BUSI"):
        pl = pl.float()
    mask_lq = lq_mask.clone()
    mask_lq[img_box.expand(mask_lq.shape) == 1] = 1
    return lq_s, pl, mask_lq, box[0]
