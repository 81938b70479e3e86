"""Bottleneck-level kernels of UNet-A at cfg2 (24x24 feature maps, 512/1024 channels) plus the 1024->512 @48 weight
gradient: the layers where wave quantisation of the persistent CTAs matters.  A/B with USTRUN_TC_TILEPLAN=0/1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch
from ustrun import engine as E, _lib as L

torch.manual_seed(0)
E.set_precision("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def act(B, H, W, C):
    a = E.Act.new(B, H, W, C); a.t.normal_(); return a


print("tile plan:", os.environ.get("USTRUN_TC_TILEPLAN", "1"))
BATCHES = [int(b) for b in os.environ.get("USTRUN_L5_B", "8,16,24").split(",")]
ONLY = os.environ.get("USTRUN_L5_ONLY", "")          # e.g. "wgrad": only that pass (for ncu captures)
ITERS = int(os.environ.get("USTRUN_BENCH_ITERS", "7"))
for B in BATCHES:
    for name, cin, cout, H in (("down4.0", 512, 1024, 24), ("down4.3", 1024, 1024, 24), ("up1.0", 1024, 512, 48)):
        x, y, g = act(B, H, H, cin), act(B, H, H, cout), act(B, H, H, cout)
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        wf, wd = E.PackedConv().get(w)
        part = torch.empty(L.MAX_PARTS * 2 * cout, device="cuda")
        fl = 2.0 * B * H * H * cin * cout * 9
        dw = torch.zeros_like(w)
        for tag, fn in (("fwd  ", lambda: E._raw_conv(x, wf, None, y, 3, part)), ("dgrad", lambda: E._raw_conv(g, wd, None, x, 3)),
                        ("wgrad", lambda: E._wgrad(g, x, dw, 1, 3))):
            if ONLY and ONLY not in tag:
                continue
            us = timeit(fn, ITERS)
            print(f"B={B:2d} {tag} {name} {cin}->{cout} @{H}   {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
        del x, y, g
    if ONLY:
        continue
    # bottleneck up-conv (ConvTranspose2d 1024->512, 24 -> 48)
    up = torch.nn.ConvTranspose2d(1024, 512, 2, stride=2).cuda()
    a = act(B, 24, 24, 1024)
    pk = E.PackedConv()
    wf, wd = pk.get(up.weight, transposed=True)
    out = act(B, 48, 48, 512)
    fl = 2.0 * B * 24 * 24 * 1024 * 512 * 4
    us = timeit(lambda: E._call("ustrun_convT2x2_fwd", L.TCGEN05, a.ptr, a.ld, E._ptr(wf), None, out.ptr, out.ld, L.BF16, B, 24, 24, 1024, 512, E._stream()))
    print(f"B={B:2d} convT fwd 1024->512 @24        {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
    us = timeit(lambda: E._call("ustrun_convT2x2_dgrad", L.TCGEN05, out.ptr, out.ld, E._ptr(wd), a.ptr, a.ld, L.BF16, B, 24, 24, 1024, 512, E._stream()))
    print(f"B={B:2d} convT dgrad 512->1024 @24      {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
