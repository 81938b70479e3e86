"""GPU parity of the individual memory-bound kernels against plain torch fp32 (through the C ABI)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import max_abs, rel_err

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # torch's fp32 reference convs must not use TF32
torch.backends.cuda.matmul.allow_tf32 = False


def _E():
    from ustrun import engine as E
    return E


def _act_from_nchw(x, dtype):
    E = _E()
    E.set_precision("fp32" if dtype == torch.float32 else "bf16")
    return E.input_nchw(x)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_roundtrip(dtype):
    E = _E()
    x = torch.randn(2, 5, 12, 20, device="cuda")
    a = _act_from_nchw(x, dtype)
    ref = x.to(dtype).float()
    assert torch.equal(a.t.float().permute(0, 3, 1, 2), ref)
    assert torch.equal(E.to_nchw(a), ref)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("ks,cin,cout,hw", [(3, 3, 16, (16, 16)), (3, 16, 24, (12, 20)), (1, 32, 16, (8, 8)), (3, 64, 64, (16, 16)),
                                          (3, 1, 64, (24, 16)), (1, 64, 2, (16, 16)), (3, 32, 4, (12, 12)), (1, 64, 4, (9, 7))])
def test_simt_conv_fwd_dgrad_wgrad(dtype, tol, ks, cin, cout, hw):
    from ustrun import _lib as L
    E = _E()
    E.set_force_simt(True)
    try:
        torch.manual_seed(0)
        x = torch.randn(2, cin, *hw, device="cuda")
        w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.2
        a = _act_from_nchw(x, dtype)
        pk = E.PackedConv()
        wf, wd = pk.get(w)
        y = a.like(cout)
        part = torch.zeros(L.MAX_PARTS * 2 * cout, device="cuda")
        nparts = E._raw_conv(a, wf, None, y, ks, part)
        xr, wr = x.to(dtype).float(), w.to(dtype).float()
        ref = F.conv2d(xr, wr, padding=ks // 2)
        assert rel_err(E.to_nchw(y), ref) < tol
        sums = part[: nparts * 2 * cout].view(nparts, 2, cout).sum(0)
        assert rel_err(sums[0], ref.sum((0, 2, 3))) < 1e-3 and rel_err(sums[1], (ref * ref).sum((0, 2, 3))) < 1e-3
        # dgrad: conv of dy with flipped/transposed weights
        dy = torch.randn_like(ref)
        g = _act_from_nchw(dy, dtype)
        gx = a.like(cin)
        E._raw_conv(g, wd, None, gx, ks)
        dyr = dy.to(dtype).float()
        ref_dx = torch.autograd.grad(F.conv2d(xr.requires_grad_(), wr, padding=ks // 2), xr, dyr)[0]
        assert rel_err(E.to_nchw(gx), ref_dx) < tol
        dw = torch.zeros_like(w)
        E._wgrad(g, a, dw, 0, ks)
        ref_dw = torch.autograd.grad(F.conv2d(xr.detach(), wr.requires_grad_(), padding=ks // 2), wr, dyr)[0]
        assert rel_err(dw, ref_dw) < max(tol, 1e-5)
        E._wgrad(g, a, dw, 1, ks)
        assert rel_err(dw, 2 * ref_dw) < max(tol, 1e-5)
    finally:
        E.set_force_simt(False)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
def test_double_conv_block_vs_torch(precision, tol):
    """conv+BN(train)+ReLU x2 forward/backward through the autograd bridge (CUDA-core kernels)."""
    from networks.unet_parts import DoubleConv
    E = _E()
    E.set_precision(precision)
    E.set_force_simt(True)
    try:
        torch.manual_seed(1)
        blk = DoubleConv(8, 16).cuda()
        ref = torch.nn.Sequential(torch.nn.Conv2d(8, 16, 3, padding=1, bias=False), torch.nn.BatchNorm2d(16), torch.nn.ReLU(),
                                  torch.nn.Conv2d(16, 16, 3, padding=1, bias=False), torch.nn.BatchNorm2d(16), torch.nn.ReLU()).cuda()
        ref.load_state_dict({k.replace("double_conv.", ""): v for k, v in blk.state_dict().items()})
        x = torch.randn(3, 8, 16, 24, device="cuda", requires_grad=True)
        x2 = x.detach().clone().requires_grad_()
        y, yr = blk(x), ref(x2)
        assert rel_err(y, yr) < tol
        gy = torch.randn_like(yr)
        y.backward(gy)
        yr.backward(gy)
        assert rel_err(x.grad, x2.grad) < tol * 3
        for (k, p), (_, q) in zip(blk.double_conv.named_parameters(), ref.named_parameters()):
            assert rel_err(p.grad, q.grad) < tol * 3, k
        for k in ("1.running_mean", "1.running_var", "4.running_mean", "4.running_var"):
            assert rel_err(blk.double_conv.state_dict()[k], ref.state_dict()[k]) < tol, k
        assert int(blk.double_conv[1].num_batches_tracked) == 1
        # eval mode uses running statistics
        blk.eval(), ref.eval()
        with torch.no_grad():
            assert rel_err(blk(x), ref(x2)) < tol
    finally:
        E.set_force_simt(False)
        E.set_precision("bf16")


@pytest.mark.parametrize("align", [False, True])
def test_upsample2x(align):
    E = _E()
    E.set_precision("fp32")
    x = torch.randn(2, 16, 6, 10, device="cuda", requires_grad=True)
    a = E.input_nchw(x.detach())
    ctx = E.Ctx(True, True)
    a.needs_grad = True
    y = E.upsample2x(ctx, a, align)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=align)
    assert max_abs(E.to_nchw(y), ref) < 1e-6
    gy = torch.randn_like(ref)
    y.g = E.input_nchw(gy)
    ctx.backward(E.GradSink())
    ref.backward(gy)
    assert max_abs(E.to_nchw(a.g), x.grad) < 1e-5
    E.set_precision("bf16")


def test_maxpool_first_max_and_skip_add():
    from ustrun import _lib as L
    E = _E()
    E.set_precision("fp32")
    y = torch.randint(0, 3, (2, 8, 8, 12), device="cuda").float().requires_grad_()   # many ties
    p = F.max_pool2d(y, 2)
    dp = torch.randn_like(p)
    p.backward(dp)
    ya, dpa, gs = E.input_nchw(y.detach()), E.input_nchw(dp), E.input_nchw(torch.ones_like(y))
    out = ya.like()
    E._call("ustrun_maxpool_bwd", ya.ptr, ya.ld, dpa.ptr, dpa.ld, gs.ptr, gs.ld, out.ptr, out.ld, L.F32, 2, 8, 12, 8, E._stream())
    assert torch.equal(E.to_nchw(out), y.grad + 1)
    E.set_precision("bf16")


def test_sgd_ema_multi_matches_torch():
    from ustrun.optim import FusedSGDEMA
    torch.manual_seed(3)
    shapes = [(64, 3, 3, 3), (64,), (5,), (128, 64, 3, 3), (2, 64, 1, 1), (2,)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ts = [torch.randn(s, device="cuda") for s in shapes]
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref_t = [t.clone() for t in ts]
    opt_ref = torch.optim.SGD(ref_p, lr=0.03, momentum=0.9, weight_decay=1e-4)
    opt = FusedSGDEMA(ps, ts, momentum=0.9, weight_decay=1e-4)
    for it in range(3):
        grads = [torch.randn_like(p) for p in ps]
        skip = it == 0                       # parameter 2 has no grad on the first step (DSBN-like)
        for i, (p, q, g) in enumerate(zip(ps, ref_p, grads)):
            if skip and i == 2:
                q.grad = None
                opt.grad_view(i).zero_()
                opt.set_has_grad(i, False)
            else:
                q.grad = g.clone()
                opt.grad_view(i).copy_(g)
                opt.set_has_grad(i, True)
        alpha = min(1 - 1 / (it + 1), 0.99)
        opt_ref.step()
        for t, q in zip(ref_t, ref_p):
            t.mul_(alpha).add_(q.data, alpha=1 - alpha)
        opt.step(lr=0.03, alpha=alpha)
        for p, q, t, u in zip(ps, ref_p, ts, ref_t):
            assert max_abs(p, q) < 1e-6 and max_abs(t, u) < 1e-6
