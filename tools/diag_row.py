"""Row-mode conv debugging: per-tap error (weights non-zero for one tap only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
from ustrun import engine as E
E.set_precision("bf16")
torch.manual_seed(0)
B, C, H, W = 1, 64, 4, 128
x = torch.randn(B, C, H, W, device="cuda")
a = E.input_nchw(x)
xr = x.bfloat16().float()
for t in range(9):
    w = torch.zeros(64, C, 3, 3, device="cuda")
    w[:, :, t // 3, t % 3] = torch.randn(64, C, device="cuda") * 0.1
    wf, _ = E.PackedConv().get(w)
    y = a.like(64)
    E._raw_conv(a, wf, None, y, 3)
    ref = F.conv2d(xr, w.bfloat16().float(), padding=1)
    got = E.to_nchw(y)
    err = float((got - ref).norm() / ref.norm())
    # is the result a column-shifted / permuted version?  try matching against shifts of the reference
    best = min(((float((got[..., 2:-2] - torch.roll(ref, s, 3)[..., 2:-2]).norm() / ref.norm()), s) for s in range(-3, 4)))
    print(f"mode {os.environ.get('USTRUN_TC_ROW')} tap {t} (dy={t//3-1}, dx={t%3-1}): rel err {err:.3e}; best column shift {best[1]} -> {best[0]:.3e}")
