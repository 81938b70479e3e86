"""Confidence bank, adaptive selection threshold and low-quality-sample bookkeeping on the device.

The reference keeps these as Python / numpy state inside ``train()`` and pays for them with device-to-host copies every step:
``train.py:754-781`` (bank FIFO ``simple_ulb / cor_pl / cor_mask / cor_hardness``, ``choice_th``), ``:612-626`` (CutMix partner
pool ``cut_img / cut_label / cut_mask`` and ``choice``), ``:741-743`` (the hardest sample ``lq_u / lq_pl / lq_mask``) and
``:722-739`` + ``obtain_all_cover_box :242-251`` (its CutMix with a labelled image).  ``ConfidenceBank`` holds the same state in
device memory -- the bank length and the threshold are device scalars -- so that a loop

    cut_img, cut_label, cut_mask = bank.pool(lb_x, lb_mask_u8)
    choice = bank.draw_choice(r_lb, r_u, perm)                       # host draws only
    out = trainer.step(dict(..., cut_img=cut_img, cut_label=cut_label, cut_mask=cut_mask, choice=choice), lq=bank.lq_input(lb_x, lb_mask_u8, new_choice))
    bank.select_lq(out["lq_idx"], ulb_w, out["pseudo_label"], out["mask"])
    bank.update(out["hardness"], ulb_w, out["pseudo_label"], out["mask"])

never synchronises with the host.  Label / mask planes are the uint8 tensors ``SSLTrainer.step`` returns.  The pool has a FIXED
size ``B_l + max_len``; entries beyond the current bank length are never referenced by ``choice``."""
from __future__ import annotations

import torch

from . import _lib as L
from .engine import _call, _ptr, _stream


class ConfidenceBank:
    def __init__(self, B_l, B_u, n_channels, H, W, label_channels=None, max_len=10, increase=1.0005, choice_th=0.1, device="cuda"):
        """label_channels: None for the softmax branch (label / mask planes [B,H,W]), C for the sigmoid branch ([B,C,H,W])."""
        L.require_device()
        if B_u > max_len:
            raise ValueError("ConfidenceBank needs queue_len >= the unlabelled batch size (the reference's default is 10 >= 4)")
        self.Bl, self.Bu, self.C, self.H, self.W, self.max_len, self.increase = B_l, B_u, n_channels, H, W, int(max_len), float(increase)
        self.lab_shape = (H, W) if label_channels is None else (label_channels, H, W)
        self.img_elems = n_channels * H * W
        self.lab_elems = H * W * (1 if label_channels is None else label_channels)
        n = B_l + self.max_len
        mk = lambda *s, dt=torch.uint8: torch.zeros(s, dtype=dt, device=device)
        self.img = [mk(n, n_channels, H, W, dt=torch.float32) for _ in range(2)]
        self.label = [mk(n, *self.lab_shape) for _ in range(2)]
        self.mask = [mk(n, *self.lab_shape) for _ in range(2)]
        for m in self.mask:
            m[:B_l] = 1                                       # cut_mask of the labelled part is all ones (train.py:614,619)
        self.hard = [mk(self.max_len, dt=torch.float64) for _ in range(2)]
        self.n = mk(1, dt=torch.int32)
        self.choice_th = torch.full((1,), float(choice_th), dtype=torch.float64, device=device)
        self._plan = mk(self.max_len + 1, dt=torch.int32)
        self.cur = 0
        self.lq_img = mk(1, n_channels, H, W, dt=torch.float32)
        self.lq_pl, self.lq_mask = mk(1, *self.lab_shape), mk(1, *self.lab_shape)
        self.has_lq = False
        self._box = mk(H, W)

    # -- CutMix partner pool and choice (train.py:612-626) --------------------------------------------
    def pool(self, lb_x, lb_mask_u8):
        """(cut_img [B_l+max_len,C,H,W] fp32, cut_label, cut_mask uint8): the labelled batch followed by the bank."""
        c = self.cur
        self.img[c][:self.Bl].copy_(lb_x)
        self.label[c][:self.Bl].copy_(lb_mask_u8.reshape((self.Bl,) + self.lab_shape))
        return self.img[c], self.label[c], self.mask[c]

    def draw_choice(self, r_lb, r_u, perm):
        """r_lb: ints in [0, B_l); r_u: floats in [0, 1); perm: a permutation of range(B_u) -- the host's random draws (any
        array-like or tensors).  Returns int32 [B_u] on the device."""
        dev = self.n.device
        a = torch.as_tensor(r_lb, dtype=torch.int32).to(dev).contiguous()
        u = torch.as_tensor(r_u, dtype=torch.float64).to(dev).contiguous()
        p = torch.as_tensor(perm, dtype=torch.int32).to(dev).contiguous()
        out = torch.empty(self.Bu, dtype=torch.int32, device=dev)
        _call("ustrun_bank_choice", _ptr(self.n), self.Bl, self.Bu, _ptr(a), _ptr(u), _ptr(p), _ptr(out), _stream())
        return out

    # -- bank FIFO + adaptive threshold (train.py:754-781) ----------------------------------------------
    def update(self, hardness, ulb_w, pseudo_label_u8, mask_u8):
        """hardness: float64 [B_u] device tensor (``ustrun.step.hardness`` / ``out['hardness']``)."""
        c, o = 1 - self.cur, self.cur
        img, pl, mk = ulb_w.float().contiguous(), pseudo_label_u8.contiguous(), mask_u8.contiguous()
        if img.shape[0] != self.Bu or pl.numel() != self.Bu * self.lab_elems or mk.numel() != self.Bu * self.lab_elems:
            raise ValueError("ConfidenceBank.update: batch tensors do not match the bank's shapes")
        off_i, off_l = self.Bl * self.img_elems * 4, self.Bl * self.lab_elems
        _call("ustrun_bank_update", _ptr(hardness.contiguous()), self.Bu, _ptr(img), _ptr(pl), _ptr(mk),
              self.img[o].data_ptr() + off_i, self.label[o].data_ptr() + off_l, self.mask[o].data_ptr() + off_l, _ptr(self.hard[o]),
              self.img[c].data_ptr() + off_i, self.label[c].data_ptr() + off_l, self.mask[c].data_ptr() + off_l, _ptr(self.hard[c]),
              _ptr(self.n), _ptr(self.choice_th), self.max_len, self.increase, self.img_elems, self.lab_elems, _ptr(self._plan), _stream())
        self.cur = c

    def state(self):
        """(n int32[1], choice_th float64[1], bank images / labels / masks / hardness of the max_len slots): device tensors."""
        c = self.cur
        return self.n, self.choice_th, self.img[c][self.Bl:], self.label[c][self.Bl:], self.mask[c][self.Bl:], self.hard[c]

    # -- hardest sample of the batch (train.py:741-743) and its CutMix (train.py:722-739) --------------
    def select_lq(self, lq_idx, ulb_w, pseudo_label_u8, mask_u8):
        _call("ustrun_lq_select", _ptr(lq_idx), _ptr(ulb_w.float().contiguous()), _ptr(pseudo_label_u8.contiguous()), _ptr(mask_u8.contiguous()),
              _ptr(self.lq_img), _ptr(self.lq_pl), _ptr(self.lq_mask), self.img_elems, self.lab_elems, _stream())
        self.has_lq = True

    def lq_box(self, lb_mask_u8, new_choice: int, fallback_box=None):
        """obtain_all_cover_box of the union {lq pseudo label, labelled mask new_choice}: uint8 [H,W] on the device."""
        hw = self.H * self.W
        lm = lb_mask_u8.reshape((self.Bl,) + self.lab_shape)[new_choice].contiguous()
        planes = [self.lq_pl.data_ptr() + i * hw for i in range(self.lab_elems // hw)] + [lm.data_ptr() + i * hw for i in range(self.lab_elems // hw)]
        planes += [None] * (4 - len(planes))
        if len(planes) > 4:
            raise NotImplementedError("cover box over more than four label planes")
        fb = None if fallback_box is None else fallback_box.to(torch.uint8).contiguous()
        _call("ustrun_cover_box", planes[0], planes[1], planes[2], planes[3], self.H, self.W, _ptr(fb), _ptr(self._box), _stream())
        self._keep = (lm, fb)
        return self._box

    def lq_input(self, lb_x, lb_mask_u8, new_choice: int, fallback_box=None):
        """The ``lq=`` argument of ``SSLTrainer.step``: (lq_u, labelled image, box) -> lq_s = lq_u (1 - box) + lb_x[new_choice] box,
        composed by the step's input kernel; None before the first ``select_lq`` (train.py:739-741: 'first')."""
        if not self.has_lq:
            return None
        return (self.lq_img, lb_x[new_choice:new_choice + 1], self.lq_box(lb_mask_u8, new_choice, fallback_box).reshape(1, self.H, self.W))
