"""Per-kernel timings of the UNet-B 16/32-channel layers (csrc/mid_conv.cu) at the cfg2b shapes (1x384x384, B=8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch
from ustrun import engine as E, _lib as L
E.set_precision("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=int(os.environ.get("USTRUN_BENCH_ITERS", "5"))):
    fn(); torch.cuda.synchronize(); ts = []
    if iters == 0:          # profiling mode (ncu): one launch per kernel
        return 1.0
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
B = 8
for name, cin, cout, H, ks in [("convd1.conv2 16->16", 16, 16, 384, 3), ("convd2.conv1 16->32", 16, 32, 192, 3), ("convd2.conv2 32->32", 32, 32, 192, 3),
                               ("convd3.conv1 32->64", 32, 64, 96, 3), ("convu1.conv1 64->32", 64, 32, 192, 3), ("convu1.conv2 32->16 1x1", 32, 16, 384, 1),
                               ("convu1.conv3 32->32", 32, 32, 384, 3)]:
    x, y, g = (E.Act.new(B, H, H, c) for c in (cin, cout, cout))
    for a in (x, y, g): a.t.normal_()
    w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05
    wf, wd = E.PackedConv().get(w)
    part = torch.empty(L.MAX_PARTS * 2 * cout, device="cuda")
    dw = torch.zeros_like(w)
    px = B * H * H
    fl = 2.0 * px * cin * cout * ks * ks
    by = px * (cin + cout) * 2
    for what, fn in (("fwd", lambda: E._raw_conv(x, wf, None, y, ks, part)), ("dgrad", lambda: E._raw_conv(g, wd, None, x, ks)), ("wgrad", lambda: E._wgrad(g, x, dw, 1, ks))):
        us = timeit(fn)
        print(f"{name:28s} @{H} {what:6s} {us:8.1f} us  {fl/us/1e6:7.1f} TFLOP/s  {by/us/1e3:7.1f} GB/s (algorithmic)", flush=True)

# logits head of UNet-B (conv3x3 32 -> classes + bias, fp32 NCHW): csrc/mid_conv.cu k_conv_mid_mma<..., HEAD>
for k in (2, 3):
    H = 384
    x = E.Act.new(B, H, H, 32); x.t.normal_()
    w = torch.randn(k, 32, 3, 3, device="cuda") * 0.05
    wf, _ = E.PackedConv().get(w, need_wd=False)
    bias = torch.zeros(k, device="cuda")
    logits = torch.empty(B, k, H, H, device="cuda")
    us = timeit(lambda: E._raw_conv(x, wf, bias, None, 3, out_nchw=logits))
    px = B * H * H
    print(f"{'out1 32->%d (head, f32 NCHW)' % k:28s} @{H} fwd    {us:8.1f} us  {2.0*px*32*k*9/us/1e6:7.1f} TFLOP/s  {px*(64+4*k)/us/1e3:7.1f} GB/s (algorithmic)", flush=True)
