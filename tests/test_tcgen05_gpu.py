"""tcgen05/TMEM/TMA conv kernels against the CUDA-core kernels and torch fp32 on the same bf16
inputs (bit-identical operands, so only the fp32 accumulation order differs)."""
import pytest
import torch
import torch.nn.functional as F

from util import max_abs, rel_err

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # torch's fp32 reference convs must not use TF32
torch.backends.cuda.matmul.allow_tf32 = False

SHAPES = [  # B, Cin, Cout, H, W
    (2, 64, 64, 16, 16),
    (1, 64, 128, 32, 32),
    (2, 128, 256, 24, 24),     # partial tiles in H
    (1, 256, 64, 18, 18),      # W not a power of two (288/16)
    (3, 64, 64, 8, 40),
    (1, 192, 320, 16, 16),     # BN=64 path with several n tiles
    (1, 64, 64, 6, 128),       # row mode (3 dx taps from one 136-pixel box), BN=64
    (2, 128, 64, 3, 384),      # row mode, two channel chunks, 3 tiles per row
    (1, 64, 128, 5, 192),      # W=192: not a multiple of 128 -> generic tiles
    (1, 128, 128, 4, 256),     # row mode BN=128
    (1, 64, 64, 4, 96),        # row-mode wgrad with 48-pixel segments
    (1, 128, 64, 4, 48),       # row-mode wgrad, swapped operands (Cout=64 < Cin)
    (1, 256, 256, 24, 24),     # row-mode wgrad, one zero-padded 32-pixel segment per row
    (8, 64, 1024, 24, 24),     # tile planner: M tiles of two 8x8 boxes (36 full tiles instead of 40 ragged ones), N=256
    (8, 128, 512, 24, 24),     # tile planner: two-box tiles and the narrower N=128 tile (all SMs get one tile)
    (5, 64, 2048, 12, 12),     # two-box tiles, 12x5 boxes (60 of 64 rows, last box of an image 2 rows), odd box count
    (7, 64, 2048, 12, 24),     # two-box tiles, 5x12 boxes ragged in W, odd box count, N=256
    (1, 64, 64, 18, 16),       # row-mode wgrad with 4 image rows per stage, last row group half empty
    (2, 64, 128, 5, 24),       # row-mode wgrad with 2 image rows per stage, odd H
]


def _setup():
    from ustrun import engine as E
    E.set_precision("bf16")
    E.set_force_simt(False)
    return E


@pytest.mark.parametrize("B,cin,cout,H,W", SHAPES)
@pytest.mark.parametrize("ks", [3, 1])
def test_tc_conv_fwd_stats_dgrad_wgrad(B, cin, cout, H, W, ks):
    from ustrun import _lib as L
    E = _setup()
    torch.manual_seed(B * 1000 + cin + cout + H)
    x = torch.randn(B, cin, H, W, device="cuda")
    w = torch.randn(cout, cin, ks, ks, device="cuda") * (1.0 / (cin * ks * ks) ** 0.5)
    a = E.input_nchw(x)
    wf, wd = E.PackedConv().get(w)
    xr, wr = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(xr, wr, padding=ks // 2)
    y = a.like(cout)
    part = torch.zeros(L.MAX_PARTS * 2 * cout, device="cuda")
    assert E._impl_for(cin, cout, L.BF16) == L.TCGEN05
    nparts = E._raw_conv(a, wf, None, y, ks, part)
    torch.cuda.synchronize()
    got = E.to_nchw(y)
    assert max_abs(got, ref.bfloat16().float()) <= 2e-2 * float(ref.abs().max()), "tcgen05 conv forward"
    assert rel_err(got, ref) < 5e-3
    sums = part[: nparts * 2 * cout].view(nparts, 2, cout).sum(0)
    assert rel_err(sums[0], ref.sum((0, 2, 3))) < 2e-3 and rel_err(sums[1], (ref * ref).sum((0, 2, 3))) < 2e-3, "fused BN statistics"
    # dgrad
    dy = torch.randn_like(ref)
    g = E.input_nchw(dy)
    dyr = dy.bfloat16().float()
    gx = a.like(cin)
    E._raw_conv(g, wd, None, gx, ks)
    ref_dx = torch.autograd.grad(F.conv2d(xr.requires_grad_(), wr, padding=ks // 2), xr, dyr)[0]
    assert rel_err(E.to_nchw(gx), ref_dx) < 5e-3, "tcgen05 dgrad"
    # wgrad (+ accumulate)
    dw = torch.zeros_like(w)
    E._wgrad(g, a, dw, 0, ks)
    ref_dw = torch.autograd.grad(F.conv2d(xr.detach(), wr.requires_grad_(), padding=ks // 2), wr, dyr)[0]
    assert rel_err(dw, ref_dw) < 2e-3, "tcgen05 wgrad"
    E._wgrad(g, a, dw, 1, ks)
    assert rel_err(dw, 2 * ref_dw) < 2e-3


@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 128, 64, 16, 16), (1, 512, 256, 12, 12), (2, 64, 64, 24, 8),
                                            (8, 1024, 512, 24, 24)])      # bottleneck up-conv: two-box tiles fwd (N=128) and dgrad (N=256)
def test_tc_conv_transpose(B, cin, cout, H, W):
    from ustrun import _lib as L
    E = _setup()
    torch.manual_seed(7)
    up = torch.nn.ConvTranspose2d(cin, cout, 2, stride=2).cuda()
    x = torch.randn(B, cin, H, W, device="cuda")
    a = E.input_nchw(x)
    a.needs_grad = True
    cat = E.Act.new(B, 2 * H, 2 * W, 2 * cout, dtype=torch.bfloat16)
    cat.t.zero_()
    out = cat.view(cout, cout)
    ctx = E.Ctx(True, True)
    E.conv_transpose2x2(ctx, a, up, out, E.PackedConv())
    xr = x.bfloat16().float().requires_grad_()
    wr = up.weight.detach().bfloat16().float().requires_grad_()
    br = up.bias.detach().clone().requires_grad_()
    ref = F.conv_transpose2d(xr, wr, br, stride=2)
    got = E.to_nchw(out)
    assert rel_err(got, ref) < 5e-3
    assert float(cat.t[..., :cout].abs().max()) == 0.0, "must not touch the skip half of the concat buffer"
    dy = torch.randn_like(ref)
    gcat = E.Act.new(B, 2 * H, 2 * W, 2 * cout, dtype=torch.bfloat16)
    gcat.t.zero_()
    gcat.t[..., cout:] = dy.permute(0, 2, 3, 1).bfloat16()
    cat.g = gcat
    sink = E.GradSink()
    ctx.backward(sink)
    ref.backward(dy.bfloat16().float())
    assert rel_err(E.to_nchw(a.g), xr.grad) < 5e-3
    assert rel_err(sink.fresh[id(up.weight)], wr.grad) < 3e-3
    assert rel_err(sink.fresh[id(up.bias)], br.grad) < 3e-3


def test_tc_matches_simt_bitwise_inputs():
    """Same bf16 operands through both implementations: differences are fp32 summation order only."""
    E = _setup()
    torch.manual_seed(11)
    x = torch.randn(2, 128, 16, 16, device="cuda")
    w = torch.randn(128, 128, 3, 3, device="cuda") * 0.03
    a = E.input_nchw(x)
    wf, _ = E.PackedConv().get(w)
    y1, y2 = a.like(128), a.like(128)
    E._raw_conv(a, wf, None, y1, 3)
    E.set_force_simt(True)
    E._raw_conv(a, wf, None, y2, 3)
    E.set_force_simt(False)
    d = (y1.t.float() - y2.t.float()).abs().max()
    assert float(d) <= 2 ** -6 * float(y2.t.float().abs().max())


MID_SHAPES = [  # B, Cin, Cout, H, W: the 16/32-channel levels of UNet-B (warp-level MMA kernels, csrc/mid_conv.cu)
    (2, 16, 16, 12, 48), (1, 16, 32, 9, 64), (2, 32, 32, 8, 32), (1, 32, 64, 6, 48), (1, 64, 32, 5, 80), (2, 32, 16, 7, 16),
    (1, 16, 16, 4, 384),
]


@pytest.mark.parametrize("B,cin,cout,H,W", MID_SHAPES)
@pytest.mark.parametrize("ks", [3, 1])
def test_mid_mma_conv_fwd_stats_dgrad_wgrad(B, cin, cout, H, W, ks):
    from ustrun import _lib as L
    E = _setup()
    torch.manual_seed(B * 100 + cin + cout + H + ks)
    x = torch.randn(B, cin, H, W, device="cuda")
    w = torch.randn(cout, cin, ks, ks, device="cuda") * (1.0 / (cin * ks * ks) ** 0.5)
    a = E.input_nchw(x)
    wf, wd = E.PackedConv().get(w)
    xr, wr = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(xr, wr, padding=ks // 2)
    y = a.like(cout)
    part = torch.zeros(L.MAX_PARTS * 2 * cout, device="cuda")
    assert E._impl_for(cin, cout, L.BF16) == L.SIMT          # dispatched inside the C ABI to the mid-channel MMA kernels
    nparts = E._raw_conv(a, wf, None, y, ks, part)
    got = E.to_nchw(y)
    assert rel_err(got, ref) < 5e-3, "mid conv forward"
    sums = part[: nparts * 2 * cout].view(nparts, 2, cout).sum(0)
    assert rel_err(sums[0], ref.sum((0, 2, 3))) < 2e-3 and rel_err(sums[1], (ref * ref).sum((0, 2, 3))) < 2e-3, "fused BN statistics"
    dy = torch.randn_like(ref)
    g = E.input_nchw(dy)
    dyr = dy.bfloat16().float()
    gx = a.like(cin)
    E._raw_conv(g, wd, None, gx, ks)
    ref_dx = torch.autograd.grad(F.conv2d(xr.requires_grad_(), wr, padding=ks // 2), xr, dyr)[0]
    assert rel_err(E.to_nchw(gx), ref_dx) < 5e-3, "mid conv dgrad"
    dw = torch.zeros_like(w)
    E._wgrad(g, a, dw, 0, ks)
    ref_dw = torch.autograd.grad(F.conv2d(xr.detach(), wr.requires_grad_(), padding=ks // 2), wr, dyr)[0]
    assert rel_err(dw, ref_dw) < 2e-3, "mid conv wgrad"
    E._wgrad(g, a, dw, 1, ks)
    assert rel_err(dw, 2 * ref_dw) < 2e-3
