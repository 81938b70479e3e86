"""Device frequency-domain style mix against the reference fixtures and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# The device kernels compute in float64.  Against the SAME formulas evaluated by numpy in float64 the float32
# outputs (ulp <= 1.2e-7 on [-1,1]) may differ by a couple of ulps (values on a float32 rounding boundary of the
# [0,255] image).  The reference fixtures were produced with numpy 2.3, whose pocketfft runs in the input precision
# (float32 here; numpy 1.x up-cast to float64): against them the difference is the reference's own single-precision
# FFT rounding, a few 1e-6 -- well inside BASELINE.json's 1e-4 floating-point tolerance.
ATOL = 3e-7
ATOL_F32_REFERENCE = 1e-5


def test_fft_mix_matches_reference_fixtures():
    from ustrun.fft_mix import amp_mix
    fx = np.load(os.path.join(GOLDEN, "fft_mix.npz"))
    for tag in sorted({k.split("/")[0] for k in fx.files}):
        src, trg = torch.from_numpy(fx[f"{tag}/mix_img"]).cuda(), torch.from_numpy(fx[f"{tag}/ulb_x_w"]).cuda()
        out = amp_mix(src, trg, fx[f"{tag}/ratio"], float(fx[f"{tag}/L"])).cpu().numpy()
        ref = fx[f"{tag}/out"]
        assert out.shape == ref.shape and out.dtype == np.float32
        err = np.abs(out - ref)
        assert err.max() <= ATOL_F32_REFERENCE, (tag, float(err.max()))
        from oracle import fft_mix_ref as Fm
        ref64 = Fm.move_transx(fx[f"{tag}/mix_img"], fx[f"{tag}/ulb_x_w"], fx[f"{tag}/ratio"].tolist(), float(fx[f"{tag}/L"]), fft_dtype=np.float64)
        assert np.abs(out - ref64).max() <= ATOL, (tag, float(np.abs(out - ref64).max()))
        assert (out == ref64).mean() > 0.98, (tag, float((out == ref64).mean()))  # almost everywhere bit-identical


@pytest.mark.parametrize("N,C,H,W,L", [(8, 1, 384, 384, 0.01), (4, 3, 256, 256, 0.01), (2, 1, 288, 288, 0.04)])
def test_fft_mix_full_size_vs_oracle(N, C, H, W, L):
    from oracle import fft_mix_ref as Fm
    from ustrun.fft_mix import amp_mix
    rng = np.random.RandomState(7)
    src = rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32)
    trg = rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32)
    ratios = rng.uniform(0, 1, N)
    ref = Fm.move_transx(src, trg, ratios.tolist(), L)
    ref64 = Fm.move_transx(src, trg, ratios.tolist(), L, fft_dtype=np.float64)
    out = amp_mix(torch.from_numpy(src).cuda(), torch.from_numpy(trg).cuda(), ratios, L).cpu().numpy()
    assert np.abs(out - ref64).max() <= ATOL
    assert np.abs(out - ref).max() <= ATOL_F32_REFERENCE
    # the mix only moves low frequencies: the result differs from the source, but stays in range
    assert out.min() >= -1.0 and out.max() <= 1.0 and np.abs(out - src).max() > 1e-3


def test_fft_mix_edge_cases():
    from ustrun.fft_mix import amp_mix
    x = torch.rand(2, 1, 64, 64, device="cuda") * 2 - 1
    y = torch.rand(2, 1, 64, 64, device="cuda") * 2 - 1
    same = amp_mix(x, y, [0.0, 0.0], 0.05)                     # ratio 0: amplitude unchanged -> identity (up to rounding)
    assert (same - ((x + 1) * 127.5).clamp(0, 255) / 127.5 + 1).abs().max() < 2e-6
    with pytest.raises(ValueError):
        amp_mix(x, y[:1], [0.5, 0.5])
    with pytest.raises(RuntimeError):
        amp_mix(x.cpu(), y.cpu(), 0.5)
    with pytest.raises(Exception):
        amp_mix(x, y, 0.5, 0.45)                                # window wider than the kernels support


def test_step_with_device_style_mix_equals_host_mix():
    """SSLTrainer.step with ``mix_ratio`` (style mix on the device) == the step fed the oracle's move_transx."""
    from networks.unet_model import UNet
    from oracle import fft_mix_ref as Fm
    from oracle import ssl_step_ref as S
    from ustrun import engine as E
    from ustrun.step import SSLTrainer
    E.set_precision("fp32")
    try:
        c, k, hw, B = 1, 2, 64, 2
        batch = S.synthetic_batch(c, k, hw, hw, B, B, seed=5)
        ratios = [0.3, 0.8]
        mix_img = batch["cut_img"][batch["choice"]]
        host_mix = torch.from_numpy(Fm.move_transx(mix_img.numpy(), batch["ulb_w"].numpy(), ratios, 0.05, fft_dtype=np.float64))
        losses = []
        for mode in ("host", "device"):
            torch.manual_seed(3)
            model, ema = UNet(c, k).cuda().train(), UNet(c, k).cuda().train()
            ema.load_state_dict(model.state_dict())
            for p_ in ema.parameters():
                p_.detach_()
            tr = SSLTrainer(model, ema, n_classes=k, threshold=0.6, fft_window=0.05)
            tr.iter_num = 3000
            b = {kk: v.cuda() for kk, v in batch.items()}
            if mode == "host":
                b["move_transx"] = host_mix.cuda()
            else:
                del b["move_transx"]
                b["mix_ratio"] = ratios
            losses.append(float(tr.step(b)["loss"]))
        assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0]), losses
    finally:
        E.set_precision("bf16")
