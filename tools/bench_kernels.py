"""Per-kernel timings at the cfg2 shapes (UNet-A, 1x384x384, B=8): CUDA events, L2 flushed between
iterations, prints a table with achieved TFLOP/s or GB/s.  Evidence for profiles/, not the product."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch
from ustrun import engine as E, _lib as L
from ustrun.step import pseudo_labels
from ustrun.loss_ops import term_forward, term_backward
from ustrun.optim import FusedSGDEMA

torch.manual_seed(0)
E.set_precision("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=int(os.environ.get('USTRUN_BENCH_ITERS', '5'))):
    fn(); torch.cuda.synchronize()
    if iters == 0:          # profiling mode (ncu): one launch per kernel
        return 1.0
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

rows = []
def report(name, us, flops=None, bytes_=None):
    rows.append(dict(kernel=name, us=round(us, 1), tflops=round(flops / us / 1e6, 1) if flops else None, gbs=round(bytes_ / us / 1e3, 1) if bytes_ else None))
    print(f"{name:44s} {us:9.1f} us " + (f"{flops/us/1e6:8.1f} TFLOP/s" if flops else "") + (f"{bytes_/us/1e3:9.1f} GB/s" if bytes_ else ""), flush=True)

def act(B, H, W, C):
    a = E.Act.new(B, H, W, C); a.t.normal_(); return a

B, H0 = 8, 384
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
# USTRUN_BENCH_STEP_SHAPE="B,H,C": shape of the pseudo-label / CE+Dice section (default: cfg2 = 8,384,2; cfg4 = 32,288,4)
STEP_SHAPE = [int(v) for v in os.environ.get("USTRUN_BENCH_STEP_SHAPE", "8,384,2").split(",")]
if ONLY == "narrow":
    layers_skip = True
else:
    layers_skip = False
# ---- tensor-core conv layers of UNet-A ----
layers = [("inc.3", 64, 64, 1), ("down1.0", 64, 128, 2), ("down1.3", 128, 128, 2), ("down2.0", 128, 256, 4), ("down2.3", 256, 256, 4),
          ("down3.0", 256, 512, 8), ("down3.3", 512, 512, 8), ("down4.0", 512, 1024, 16), ("down4.3", 1024, 1024, 16),
          ("up1.0", 1024, 512, 8), ("up2.0", 512, 256, 4), ("up3.0", 256, 128, 2), ("up4.0", 128, 64, 1)]
_sel = os.environ.get("USTRUN_BENCH_LAYERS", "")
if _sel:
    layers = [l for l in layers if l[0] in _sel.split(",")]
for name, cin, cout, d in ([] if layers_skip else layers):
    H = H0 // d
    x, y, g = act(B, H, H, cin), act(B, H, H, cout), act(B, H, H, cout)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, wd = E.PackedConv().get(w)
    part = torch.empty(L.MAX_PARTS * 2 * cout, device="cuda")
    fl = 2.0 * B * H * H * cin * cout * 9
    report(f"tc_conv fwd  {name} {cin}->{cout} @{H}", timeit(lambda: E._raw_conv(x, wf, None, y, 3, part)), flops=fl)
    report(f"tc_conv dgrad {name} {cout}->{cin} @{H}", timeit(lambda: E._raw_conv(g, wd, None, x, 3)), flops=fl)
    dw = torch.zeros_like(w)
    report(f"tc_wgrad     {name} @{H}", timeit(lambda: E._wgrad(g, x, dw, 1, 3)), flops=fl)
    del x, y, g
# ---- narrow layers ----
H = H0
x1, y64 = act(B, H, H, 1), act(B, H, H, 64)
w1 = torch.randn(64, 1, 3, 3, device="cuda"); wf1, wd1 = E.PackedConv().get(w1)
part = torch.empty(L.MAX_PARTS * 2 * 64, device="cuda")
npx = B * H * H
report("narrow_in first conv 1->64", timeit(lambda: E._raw_conv(x1, wf1, None, y64, 3, part)), bytes_=npx * (2 + 128))
dw1 = torch.zeros_like(w1)
report("narrow wgrad first conv", timeit(lambda: E._wgrad(y64, x1, dw1, 1, 3)), bytes_=npx * (2 + 128))
wh = torch.randn(2, 64, 1, 1, device="cuda"); wfh, wdh = E.PackedConv().get(wh)
logits = torch.empty(B, 2, H, H, device="cuda")
bias = torch.zeros(2, device="cuda")
report("narrow_out head 64->2 (NCHW f32)", timeit(lambda: E._raw_conv(y64, wfh, bias, None, 1, out_nchw=logits)), bytes_=npx * (128 + 8))
g2 = act(B, H, H, 2)
np_ = ctypes.c_int(0)
report("narrow_in head dgrad 2->64", timeit(lambda: E._call("ustrun_conv_fwd", L.SIMT, g2.ptr, g2.ld, E._ptr(wdh), None, y64.ptr, y64.ld, L.BF16, B, H, H, 2, 64, 1, 0, None, ctypes.byref(np_), E._stream())), bytes_=npx * (4 + 128))
dwh = torch.zeros_like(wh)
nb = L.lib.ustrun_conv_wgrad_workspace_bytes(L.SIMT, B, H, H, 64, 2, 1)
ws = torch.empty(int(nb), dtype=torch.uint8, device="cuda")
report("narrow wgrad head", timeit(lambda: E._call("ustrun_conv_wgrad", L.SIMT, g2.ptr, g2.ld, y64.ptr, y64.ld, E._ptr(dwh), 1, L.BF16, B, H, H, 64, 2, 1, E._ptr(ws), int(nb), E._stream())), bytes_=npx * (4 + 128))
# ---- BN / pool elementwise at level 1 ----
raw, y, G, pooled = act(B, H, H, 64), act(B, H, H, 64), act(B, H, H, 64), act(B, H // 2, H // 2, 64)
st = torch.rand(7 * 64, device="cuda") + 0.5
scale, shift, mean, rstd, coef = st[:64], st[64:128], st[128:192], st[192:256], st[256:448]
S = E._stream
nb16 = npx * 64 * 2
report("bn_act (BN+ReLU apply)", timeit(lambda: E._call("ustrun_bn_act_fwd", raw.ptr, 64, E._ptr(scale), E._ptr(shift), 1, y.ptr, 64, None, 0, L.BF16, B, H, H, 64, S())), bytes_=2 * nb16)
report("bn_act_pool (BN+ReLU+maxpool)", timeit(lambda: E._call("ustrun_bn_act_fwd", raw.ptr, 64, E._ptr(scale), E._ptr(shift), 1, y.ptr, 64, pooled.ptr, 64, L.BF16, B, H, H, 64, S())), bytes_=2.25 * nb16)
part = torch.empty(L.MAX_PARTS * 128, device="cuda")
report("bn_bwd_reduce", timeit(lambda: E._call("ustrun_bn_bwd_reduce", G.ptr, 64, raw.ptr, 64, E._ptr(mean), E._ptr(rstd), E._ptr(scale), E._ptr(shift), 1, L.BF16, npx, 64, E._ptr(part), ctypes.byref(np_), S())), bytes_=2 * nb16)
report("bn_bwd_apply", timeit(lambda: E._call("ustrun_bn_bwd_apply", G.ptr, 64, raw.ptr, 64, E._ptr(mean), E._ptr(rstd), E._ptr(scale), E._ptr(shift), E._ptr(coef), 1, y.ptr, 64, L.BF16, npx, 64, S())), bytes_=3 * nb16)
report("maxpool_bwd (+skip add)", timeit(lambda: E._call("ustrun_maxpool_bwd", y.ptr, 64, pooled.ptr, 64, G.ptr, 64, raw.ptr, 64, L.BF16, B, H, H, 64, S())), bytes_=3.25 * nb16)
sc4 = torch.empty(4 * 64, device="cuda"); pr = torch.rand(148 * 128, device="cuda")
gam = torch.ones(64, device="cuda"); rm = torch.zeros(64, device="cuda"); rv = torch.ones(64, device="cuda"); nbt = torch.zeros(1, dtype=torch.int64, device="cuda")
report("bn_finalize (148 partial rows)", timeit(lambda: E._call("ustrun_bn_finalize", E._ptr(pr), 148, 64, float(npx), E._ptr(gam), E._ptr(rm), None, E._ptr(rm), E._ptr(rv), E._ptr(nbt), 0.1, 1e-5, 1, E._ptr(sc4[:64]), E._ptr(sc4[64:128]), E._ptr(sc4[128:192]), E._ptr(sc4[192:]), None, S())))
# ---- step kernels ----
B, H, C = STEP_SHAPE
npx = B * H * H
t = [torch.randn(B, C, H, H, device="cuda") * 3 for _ in range(4)]
box = (torch.rand(B, H, H, device="cuda") > 0.7).to(torch.uint8)
cl = torch.randint(0, C, (B, H, H), device="cuda", dtype=torch.uint8); cm = torch.ones(B, H, H, device="cuda", dtype=torch.uint8)
ch = torch.arange(B, device="cuda", dtype=torch.int32)
report("pseudo_label_softmax (fused, 9 planes)", timeit(lambda: pseudo_labels(t[0], t[1], t[2], box, cl, cm, ch, 0.95, "softmax", student_logits=t[3])), bytes_=npx * (4 * C * 4 + 3 + 9))
coefs = {}
def fwd():
    coefs["l"], coefs["c"] = term_forward(t[0], cl, cm, "softmax")
report("ce_dice pass1 + finalize", timeit(fwd), bytes_=npx * (C * 4 + 2))
dl = torch.empty_like(t[0])
report("ce_dice pass2 (gradient)", timeit(lambda: term_backward(t[0], cl, cm, "softmax", coefs["c"], out=dl)), bytes_=npx * (2 * C * 4 + 2))
from networks.unet_model import UNet
m, e = UNet(1, 2).cuda(), UNet(1, 2).cuda()
opt = FusedSGDEMA(list(m.parameters()), list(e.parameters()))
opt.step(0.03, 0.99)
nparam = sum(p.numel() for p in m.parameters())
report("sgd_ema_multi (31.0M params, 1 launch)", timeit(lambda: opt.step(0.03, 0.99)), bytes_=nparam * 28)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
