"""On-device hardness of the unlabelled batch -- integer counts + float64: BIT-EXACT against the
reference fixtures and against the oracle at full size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("mode", ["binary", "2label", "3label"])
def test_hardness_matches_reference_fixture_bit_exact(mode):
    from ustrun.step import hardness
    fx = np.load(os.path.join(GOLDEN, "hardness.npz"))
    stu, tea = torch.from_numpy(fx[f"{mode}/stu"]).cuda(), torch.from_numpy(fx[f"{mode}/tea"]).cuda()
    h, lq, _ = hardness(stu, tea, mode)
    assert np.array_equal(h.cpu().numpy(), fx[f"{mode}/hardness"])
    assert int(lq.item()) == int(fx[f"{mode}/lq_idx"])
    h1, lq1, _ = hardness(stu, tea, mode, first_epoch=True)                 # train.py:711-713
    assert bool((h1 == 1).all()) and int(lq1.item()) == 0


@pytest.mark.parametrize("mode,shape,hi", [("binary", (8, 384, 384), 2), ("2label", (4, 2, 256, 256), 2), ("3label", (32, 288, 288), 4)])
def test_hardness_full_size_vs_oracle(mode, shape, hi):
    from oracle import hardness_ref as Hr
    from ustrun.step import hardness
    rng = np.random.RandomState(3)
    stu = rng.randint(0, hi, shape).astype(np.uint8)
    tea = np.where(rng.rand(*shape) < 0.9, stu, rng.randint(0, hi, shape)).astype(np.uint8)
    tea[2] = 0
    ref_h, ref_lq, ref_d = Hr.hardness(stu, tea, mode)
    h, lq, d = hardness(torch.from_numpy(stu).cuda(), torch.from_numpy(tea).cuda(), mode)
    assert np.array_equal(h.cpu().numpy(), ref_h) and np.array_equal(d.cpu().numpy(), ref_d)
    assert int(lq.item()) == ref_lq


def test_hardness_from_step_planes():
    """The planes the fused pseudo-label kernel already writes feed the hardness kernel directly."""
    from oracle import hardness_ref as Hr
    from ustrun.step import hardness, pseudo_labels
    torch.manual_seed(0)
    B, C, H, W = 4, 2, 64, 64
    t = [torch.randn(B, C, H, W, device="cuda") * 3 for _ in range(4)]
    box = (torch.rand(B, H, W, device="cuda") > 0.7).float()
    cl = torch.randint(0, C, (B, H, W), device="cuda")
    cm = torch.ones(B, H, W, device="cuda")
    comp = pseudo_labels(t[0], t[1], t[2], box, cl, cm, torch.arange(B, device="cuda"), 0.6, "softmax", student_logits=t[3])
    h, lq, _ = hardness(comp["stu_pseudo_label"], comp["pseudo_label"], "binary")
    ref_h, ref_lq, _ = Hr.hardness(comp["stu_pseudo_label"].cpu().numpy(), comp["pseudo_label"].cpu().numpy(), "binary")
    assert np.array_equal(h.cpu().numpy(), ref_h) and int(lq.item()) == ref_lq
    with pytest.raises(ValueError):
        hardness(comp["stu_pseudo_label"], comp["pseudo_label"][:2], "binary")
