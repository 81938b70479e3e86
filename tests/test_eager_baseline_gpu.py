"""The "kernel to beat" (SURVEY 8d): the reference's step executed by PyTorch eager / cuDNN on the SAME
B200 -- here through the oracle restatement (pinned bit-for-bit to the reference modules) moved to CUDA, in
fp32 (TF32 off), bf16 autocast and fp16 autocast (what the reference ships, train.py:30,551).  The oracle is
only the checker / yardstick: the test asserts loss agreement with the sm_100a step at the BASELINE.json
configs[1] size and records the timings in gpurun_out/eager_baseline.json."""
import json
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time_eager(fwd, st_s, st_t, batch, k, autocast_dtype, steps=3):
    from oracle import ssl_step_ref as S
    bufs = {}
    ts, loss = [], None
    for i in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
            out = S.ssl_step(fwd, st_s, st_t, bufs, batch, n_classes=k, iter_num=30000 + i, max_iterations=60000, lq=batch["ulb_w"][:1])
        torch.cuda.synchronize()
        if i == 0:
            loss = float(out["loss"])
        else:
            ts.append(time.perf_counter() - t0)
    return min(ts), loss


def test_eager_cudnn_step_vs_sm100a_step():
    from networks.unet_model import UNet
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import engine as E
    from ustrun.step import SSLTrainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    c, k, H, W, B = 1, 2, 384, 384, 8
    batch = {kk: v.cuda() for kk, v in S.synthetic_batch(c, k, H, W, B, B, seed=1337).items()}
    fwd = lambda s, x: U.unet_a_forward(s, x, True)
    res = {}
    for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16), ("fp16_autocast", torch.float16)):
        st_s = {n: v.cuda() for n, v in U.init_unet_a(c, k, seed=1337).items()}
        st_t = {n: v.clone() for n, v in st_s.items()}
        sec, loss = _time_eager(fwd, st_s, st_t, batch, k, dt)
        res[name] = {"ms_per_step": sec * 1e3, "images_per_s": 2 * B / sec, "loss_step0": loss}
        del st_s, st_t
        torch.cuda.empty_cache()
    # ours, same inputs and initial weights
    E.set_precision("bf16")
    model, ema = UNet(c, k), UNet(c, k)
    st = U.init_unet_a(c, k, seed=1337)
    model.load_state_dict(st), ema.load_state_dict(st)
    model, ema = model.cuda().train(), ema.cuda().train()
    for p in ema.parameters():
        p.detach_()
    E.reserve_pool(24 << 30)                 # steady-state timing: no cudaMalloc growth during the timed steps
    tr = SSLTrainer(model, ema, n_classes=k, max_iterations=60000)
    tr.iter_num = 30000
    dev = dict(batch)
    for kk in ("lb_mask", "cut_label", "cut_mask", "box"):
        dev[kk] = dev[kk].to(torch.uint8)
    lq = dev["ulb_w"][:1].contiguous()
    out0 = tr.step(dev, lq=lq)
    loss0 = float(out0["loss"])
    for _ in range(3):
        tr.step(dev, lq=lq)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        tr.step(dev, lq=lq)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / 5
    res["ustrun_sm100a_bf16"] = {"ms_per_step": sec * 1e3, "images_per_s": 2 * B / sec, "loss_step0": loss0}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "eager_baseline.json"), "w"), indent=1)
    print(json.dumps(res))
    ref = res["fp32"]["loss_step0"]
    assert abs(loss0 - ref) < 1e-2 * abs(ref), (loss0, ref)                      # BASELINE.json: losses within 1e-2 (bf16)
    assert abs(res["bf16_autocast"]["loss_step0"] - ref) < 1e-2 * abs(ref)
    assert res["ustrun_sm100a_bf16"]["ms_per_step"] < res["bf16_autocast"]["ms_per_step"], "slower than torch eager bf16"
