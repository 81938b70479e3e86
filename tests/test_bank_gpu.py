"""Confidence bank / adaptive threshold / CutMix partner choice / low-quality sample on the device (``ustrun.bank``) against
tests/golden/bank.npz -- produced by executing the reference's own lines (train.py:754-781, :612-626, :722-739, :242-251;
oracle/make_golden.py::case_bank) on a scripted sequence of eight steps per dataset that walks through every branch: empty bank,
first inserts, FIFO overflow, no insert (threshold relaxes), threshold tightening.  Integer / byte / float64 bookkeeping:
bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ds,C,lab_c", [("prostate", 1, None), ("fundus", 3, 2)])
def test_bank_sequence_matches_reference_lines(ds, C, lab_c):
    from ustrun.bank import ConfidenceBank
    fx = np.load(os.path.join(GOLDEN, "bank.npz"))
    Bl = Bu = 4
    H = W = 16
    bank = ConfidenceBank(Bl, Bu, C, H, W, label_channels=lab_c, max_len=6, increase=1.0005)
    t = lambda k: torch.from_numpy(fx[k]).cuda()
    L_ = H * W * (lab_c or 1)
    n_prev = 0
    for step in range(8):
        p = f"{ds}/{step}"
        lb_x, lb_mask, ulb_w = t(p + "/lb_x_w"), t(p + "/lb_mask"), t(p + "/ulb_x_w")
        pl, mask = t(p + "/pseudo_label"), t(p + "/mask")
        # pool + choice with the bank of the previous step
        cut_img, cut_label, cut_mask = bank.pool(lb_x, lb_mask)
        choice = bank.draw_choice(fx[p + "/r_lb"], fx[p + "/r_u"], fx[p + "/perm"])
        assert np.array_equal(choice.cpu().numpy(), fx[p + "/choice"]), (step, choice, fx[p + "/choice"])
        assert int(choice.max()) < Bl + max(n_prev, 1) or n_prev == 0
        assert torch.equal(cut_img[:Bl], lb_x) and bool((cut_mask[:Bl] == 1).all())
        if step > 0:
            q = f"{ds}/{step - 1}"
            n = int(fx[q + "/n_bank_after"])
            assert torch.equal(cut_img[Bl:Bl + n], t(q + "/bank_img")) and torch.equal(cut_label[Bl:Bl + n].reshape(n, L_), t(q + "/bank_pl").reshape(n, L_))
            assert torch.equal(cut_mask[Bl:Bl + n].reshape(n, L_), t(q + "/bank_mask").reshape(n, L_))
        # low-quality sample of this batch and its CutMix box / image
        if (p + "/lq_s") in fx.files:
            lq_idx = torch.tensor([int(fx[p + "/lq_idx"])], dtype=torch.int32, device="cuda")
            bank.select_lq(lq_idx, ulb_w, pl, mask.reshape(pl.shape) if lab_c is None else mask)
            nc = int(fx[p + "/lq_new_choice"])
            lq_u, lb_img, box = bank.lq_input(lb_x, lb_mask, nc)
            assert np.array_equal(box[0].cpu().numpy(), fx[p + "/lq_box"])
            lq_s = torch.where(box.bool().unsqueeze(1), lb_img, lq_u)
            assert np.array_equal(lq_s.cpu().numpy(), fx[p + "/lq_s"])
        # update
        hard = t(p + "/hardness")
        bank.update(hard, ulb_w, pl, mask.reshape(pl.shape) if lab_c is None else mask)
        n_dev, th, b_img, b_pl, b_mask, b_hard = bank.state()
        n = int(fx[p + "/n_bank_after"])
        assert int(n_dev) == n, (step, int(n_dev), n)
        assert float(th) == float(fx[p + "/choice_th_after"]), (step, float(th), float(fx[p + "/choice_th_after"]))
        assert torch.equal(b_img[:n], t(p + "/bank_img"))
        assert torch.equal(b_pl[:n].reshape(n, L_), t(p + "/bank_pl").reshape(n, L_)) and torch.equal(b_mask[:n].reshape(n, L_), t(p + "/bank_mask").reshape(n, L_))
        assert np.array_equal(b_hard[:n].cpu().numpy(), fx[p + "/bank_hardness"])
        n_prev = n


def test_cover_box_empty_region_uses_fallback():
    from ustrun.bank import ConfidenceBank
    bank = ConfidenceBank(2, 2, 1, 16, 16, max_len=4)
    z = torch.zeros(2, 16, 16, dtype=torch.uint8, device="cuda")
    bank.select_lq(torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(2, 1, 16, 16, device="cuda"), z, z)
    fb = torch.zeros(16, 16, dtype=torch.uint8, device="cuda")
    fb[3:9, 2:5] = 1
    assert torch.equal(bank.lq_box(z, 1, fallback_box=fb), fb)
    assert int(bank.lq_box(z, 1).sum()) == 0


def test_step_loop_with_bank_has_no_host_dependency():
    """A few iterations of the loop sketched in ustrun/bank.py: hardness, lq_idx, bank and choice stay on the device; the
    bank length is read back only HERE, to check that the loop filled it."""
    from networks.unet_model import UNet
    from ustrun import synth as S
    from ustrun.bank import ConfidenceBank
    from ustrun.step import SSLTrainer
    torch.manual_seed(1337)
    s, t = UNet(1, 2), UNet(1, 2)
    t.load_state_dict(s.state_dict())
    for p in t.parameters():
        p.detach_()
    tr = SSLTrainer(s.cuda().train(), t.cuda().train(), n_classes=2, threshold=0.6, hardness_mode="binary", use_graph=True, lanes=2)
    bank = ConfidenceBank(2, 2, 1, 64, 64, max_len=4, choice_th=0.9)
    rng = np.random.RandomState(0)
    for i in range(6):
        b = {k: v.cuda() for k, v in S.synthetic_batch(1, 2, 64, 64, 2, 2, seed=300 + i).items()}
        lb_mask = b["lb_mask"].to(torch.uint8)
        cut_img, cut_label, cut_mask = bank.pool(b["lb_x"], lb_mask)
        choice = bank.draw_choice(rng.randint(0, 2, 2), rng.uniform(0, 1, 2), rng.permutation(2))
        batch = dict(b, cut_img=cut_img, cut_label=cut_label, cut_mask=cut_mask, choice=choice, lb_mask=lb_mask)
        out = tr.step(batch, lq=bank.lq_input(b["lb_x"], lb_mask, int(rng.randint(0, 2))))
        bank.select_lq(out["lq_idx"], b["ulb_w"], out["pseudo_label"], out["mask"])
        bank.update(out["hardness"], b["ulb_w"], out["pseudo_label"], out["mask"])
    n, th, *_ = bank.state()
    assert 0 < int(n) <= 4 and 0.0 < float(th) <= 0.9
    assert torch.isfinite(out["loss"])
