"""Input pipeline on the device: Normalize_tf + ToTensor (dataloaders/custom_transforms.py:650-684, 728-753)
as ``ustrun.step.normalize_u8`` -- bit-exact against tests/golden/inputs.npz, the output of the reference's own two classes --
and the SSL step fed with the uint8 H x W x C batches the loaders hold BEFORE that step (4x fewer bytes over PCIe), which must
equal the step fed with the normalised float32 tensors bit for bit."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_normalize_u8_is_bit_exact():
    from ustrun.step import normalize_u8
    fx = np.load(os.path.join(GOLDEN, "inputs.npz"))
    for tag in ("rgb", "gray"):
        got = normalize_u8(torch.from_numpy(fx[tag + "/u8"]).cuda())
        assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), fx[tag + "/out"]), tag


@pytest.mark.parametrize("device_mix", [False, True])
def test_step_on_uint8_batches_equals_step_on_normalised_batches(device_mix):
    from networks.unet_model import UNet
    from ustrun import synth as S
    from ustrun.step import SSLTrainer, normalize_u8

    def pair():
        torch.manual_seed(1337)
        s, t = UNet(3, 2), UNet(3, 2)
        t.load_state_dict(s.state_dict())
        for p in t.parameters():
            p.detach_()
        return s.cuda().train(), t.cuda().train()

    g = torch.Generator().manual_seed(9)
    u8 = {k: torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8).cuda() for k in ("lb_x", "ulb_w", "ulb_s", "cut_img")}
    u8["cut_img"] = u8["lb_x"].clone()
    rest = {k: v.cuda() for k, v in S.synthetic_batch(3, 2, 64, 64, 2, 2, seed=5).items() if k not in u8}
    if device_mix:
        del rest["move_transx"]
        rest["mix_ratio"] = [0.25, 0.6]
    (s0, t0), (s1, t1) = pair(), pair()
    tr0, tr1 = SSLTrainer(s0, t0, n_classes=2, threshold=0.6), SSLTrainer(s1, t1, n_classes=2, threshold=0.6, use_graph=True, lanes=2)
    for _ in range(4):
        a = tr0.step({**rest, **{k: normalize_u8(v) for k, v in u8.items()}}, lq=normalize_u8(u8["ulb_w"][:1]))
        b = tr1.step({**rest, **u8}, lq=u8["ulb_w"][:1].contiguous())
        assert torch.equal(a["loss"], b["loss"])
    for p, q in zip(s0.parameters(), s1.parameters()):
        assert torch.equal(p, q)
