"""Tier A: the reference's training loop body written with plain torch ops, torch.optim.SGD, GradScaler and
autocast -- exactly the statements of train.py:638-851 -- driving the DROP-IN modules (``networks.unet_model.UNet``,
``utils.losses.DiceLossWithMask`` resolved from ust-run_b200/) through ``model(x)`` / ``loss.backward()`` /
``optimizer.step()``.  Checked against the step fixture produced by the reference's own modules."""
import os

import numpy as np
import pytest
import torch
from torch.nn import CrossEntropyLoss

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest(v):
    v = v.detach().double().flatten().cpu()
    return np.concatenate([[v.sum().item(), v.abs().sum().item()], v[:4].numpy()]).astype(np.float64)


@pytest.mark.parametrize("amp", [False, True], ids=["amp0", "amp1"])
def test_reference_loop_body_with_dropin_modules(amp):
    from networks.unet_model import UNet                      # ust-run_b200/networks (precedes the reference on sys.path)
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import engine as E
    from utils import losses
    fx = np.load(os.path.join(GOLDEN, "step_a_softmax_c1_k2_32_b2_it0_bank0.npz"))
    c, k, hw, B, iter_num, max_iterations, base_lr, thr = 1, 2, 32, 2, 0, 30000, 0.03, float(fx["threshold"])
    torch.manual_seed(1337)
    st_s, st_t = U.init_unet_a(c, k), U.init_unet_a(c, k)
    model, ema_model = UNet(n_channels=c, n_classes=k), UNet(n_channels=c, n_classes=k)
    model.load_state_dict(st_s), ema_model.load_state_dict(st_t)
    model, ema_model = model.cuda(), ema_model.cuda()
    for p in ema_model.parameters():                           # train.py:501-502
        p.detach_()
    optimizer = torch.optim.SGD(model.parameters(), lr=base_lr, momentum=0.9, weight_decay=0.0001)     # train.py:512
    ce_loss, dice_loss = CrossEntropyLoss(reduction="none"), losses.DiceLossWithMask(k)                # train.py:518-521
    scaler = torch.amp.GradScaler("cuda", enabled=amp)                                                 # train.py:551
    b = {kk: v.cuda() for kk, v in S.synthetic_batch(c, k, hw, hw, B, B, seed=1337).items()}
    E.set_precision("fp32")
    try:
        model.train(), ema_model.train()                       # train.py:565-566
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):                                 # train.py:611
            img_box = b["box"].unsqueeze(1)
            mix_img = b["cut_img"][b["choice"]]
            with torch.no_grad():                              # train.py:638-667
                t1 = ema_model(b["ulb_w"])
                t2 = ema_model(b["ulb_w"] * (1 - img_box) + mix_img * img_box)
                t3 = ema_model(mix_img * (1 - img_box) + b["ulb_w"] * img_box)
                comp = S.compose(t1.float(), t2.float(), t3.float(), b["box"], b["cut_label"], b["cut_mask"], b["choice"], thr, "softmax")
            model(b["ulb_w"])                                  # train.py:668 (BN running statistics only)
            outs = [model(b["lb_x"]),                          # train.py:699-702
                    model(b["ulb_s"] * (1 - img_box) + b["move_transx"] * img_box),
                    model(b["move_transx"] * (1 - img_box) + b["ulb_s"] * img_box),
                    model(b["ulb_s"])]
            tg = [(b["lb_mask"], None), (comp["pseudo_label_ul"], comp["mask_ul"]), (comp["pseudo_label_lu"], comp["mask_lu"]),
                  (comp["pseudo_label_w"], comp["mask_w"])]
            terms = []
            for o, (t, m) in zip(outs, tg):                    # train.py:816-836
                cel = ce_loss(o, t)
                if m is not None:
                    cel = cel * m.squeeze(1)
                terms.append(cel.mean() + dice_loss(o, t.unsqueeze(1), mask=m, softmax=True))
            cw = S.consistency_weight(iter_num, max_iterations)
            loss = terms[0] + cw * (terms[1] + terms[2] + cw * terms[3])                               # train.py:838
        optimizer.zero_grad()                                  # train.py:840-848
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        alpha = min(1 - 1 / (iter_num + 1), 0.99)              # train.py:87-93
        for ema_param, param in zip(ema_model.parameters(), model.parameters()):
            ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)
        torch.cuda.synchronize()
    finally:
        E.set_precision("bf16")
    assert abs(float(loss) - float(fx["loss"])) <= 1e-4 * abs(float(fx["loss"])), (float(loss), float(fx["loss"]))
    assert np.allclose([float(t) for t in terms], fx["terms"], rtol=1e-4, atol=1e-6)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    # BN running statistics after 5 student / 3 teacher forwards, and the EMA copy (alpha = 0 at step 0: teacher := student)
    for mod, tag in ((model, "student_after/"), (ema_model, "teacher_after/")):
        for name, v in mod.state_dict().items():
            ref = fx[tag + name]
            if name.endswith("num_batches_tracked"):
                assert _digest(v)[0] == ref[0], name
            elif "running_" in name:
                got = _digest(v)
                assert abs(got[1] - ref[1]) <= 2e-4 * abs(ref[1]) and np.all(np.abs(got[2:] - ref[2:]) <= 2e-4 * max(np.abs(ref[2:]).max(), 1e-3)), name
    for (n_s, p_s), (n_t, p_t) in zip(model.named_parameters(), ema_model.named_parameters()):
        assert torch.equal(p_s.data, p_t.data), n_s            # alpha == 0 copies the student into the teacher
