"""Does a block of an HBM-bound BatchNorm kernel run on the SAME SM next to a tcgen05 conv CTA?  (multi-lane step)

Times N launches of a tensor-bound conv layer (stream A) and N launches of an HBM-bound elementwise pass (stream B) separately
and concurrently.  If blocks co-reside, t(both) approaches max(t_A, t_B); if the conv CTAs own their SMs, it is t_A + t_B.

    python tools/coreside_probe.py            # USTRUN_TC_COSHARE=0|1 selects the conv kernel's shared-memory headroom
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ust-run_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
import torch.nn as nn

from ustrun import _lib as L
from ustrun import engine as E


def main():
    torch.manual_seed(0)
    dev = "cuda"
    N = 20
    results = {}
    for name, (B, H, W, cin, cout) in {"256->256 @96 (N=256 tile)": (8, 96, 96, 256, 256), "128->128 @192 (N=128 tile)": (8, 192, 192, 128, 128),
                                        "64->64 @384 (row mode)": (8, 384, 384, 64, 64)}.items():
        conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False).to(dev)
        pk = E.PackedConv()
        wf, _ = pk.get(conv.weight, need_wd=False)
        x = E.Act(torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16))
        y = x.like(cout)
        part = torch.empty(L.MAX_PARTS * 2 * cout, dtype=torch.float32, device=dev)
        # HBM-bound pass on a level-1 sized tensor (8 x 384 x 384 x 64 bf16 = 151 MB in, 151 MB out)
        raw = E.Act(torch.randn(8, 384, 384, 64, device=dev).to(torch.bfloat16))
        out = raw.like()
        sc = torch.ones(64, device=dev)
        sh = torch.zeros(64, device=dev)
        g = raw.like()
        coef = torch.ones(3 * 64, device=dev)

        def conv_fn():
            E._raw_conv(x, wf, None, y, 3, part)

        def bn_act_fn():
            E._call("ustrun_bn_act_fwd", raw.ptr, raw.ld, E._ptr(sc), E._ptr(sh), L.ACT_RELU, out.ptr, out.ld, None, 0, raw.dtype_code, raw.B, raw.H, raw.W, 64, E._stream())

        def bn_apply_fn():
            E._call("ustrun_bn_bwd_apply", g.ptr, g.ld, raw.ptr, raw.ld, E._ptr(sh), E._ptr(sc), E._ptr(sc), E._ptr(sh), E._ptr(coef), L.ACT_RELU, out.ptr, out.ld,
                    raw.dtype_code, raw.npix, 64, E._stream())

        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fa, fb):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
            if fa:
                with torch.cuda.stream(sa):
                    for _ in range(N):
                        fa()
            if fb:
                with torch.cuda.stream(sb):
                    for _ in range(N):
                        fb()
            torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / N * 1e3

        for _ in range(2):
            timed(conv_fn, bn_act_fn)
        ta = timed(conv_fn, None)
        for bname, bf in (("bn_act", bn_act_fn), ("bn_bwd_apply", bn_apply_fn)):
            tb = timed(None, bf)
            tab = timed(conv_fn, bf)
            print(f"{name:28s} conv {ta:7.1f} us | {bname:12s} {tb:7.1f} us | both {tab:7.1f} us  (sum {ta + tb:7.1f}, max {max(ta, tb):7.1f})  overlap {(ta + tb - tab) / min(ta, tb):5.2f}", flush=True)


if __name__ == "__main__":
    print("USTRUN_TC_COSHARE =", os.environ.get("USTRUN_TC_COSHARE", "(default 1)"))
    main()
