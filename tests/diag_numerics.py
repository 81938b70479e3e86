"""Numerics diagnostics (GPU): where do fp32/bf16 errors come from?  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import numpy as np, torch
from oracle import unet_ref as U, ssl_step_ref as S
from ustrun import engine as E
from utils.losses import MaskedCEDice

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

def oracle(st, x, tgt, msk, k, fwd, dtype):
    st = {kk: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for kk, v in st.items()}
    params, _ = U.split_state(st)
    for p in params.values(): p.requires_grad_(True)
    logits = fwd(st, x.to(dtype))
    loss = S.masked_term(logits, tgt, msk.to(dtype), k, "softmax")
    loss.backward()
    return logits.detach(), loss.detach(), {kk: p.grad for kk, p in params.items()}

def ours(make, st, x, tgt, msk, k, precision, simt):
    E.set_precision(precision); E.set_force_simt(simt)
    mod = make(); mod.load_state_dict(st); mod = mod.cuda().train()
    logits = mod(x.cuda()); loss = MaskedCEDice(k)(logits, tgt.cuda(), msk.cuda()); loss.backward()
    return logits.detach(), loss.detach(), {kk: p.grad for kk, p in mod.named_parameters()}

def autocast_ref(st, x, tgt, msk, k, fwd):
    st = {kk: v.clone().cuda() for kk, v in st.items()}
    params, _ = U.split_state(st)
    for p in params.values(): p.requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = fwd(st, x.cuda())
    loss = S.masked_term(logits.float(), tgt.cuda(), msk.cuda(), k, "softmax")
    loss.backward()
    return logits.detach().float(), loss.detach(), {kk: p.grad for kk, p in params.items()}

def case(name, make, init, fwd, c, k, hw, B):
    torch.manual_seed(0)
    st = init()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, c, hw, hw, generator=g) * 2 - 1
    tgt = torch.randint(0, k, (B, hw, hw), generator=g)
    msk = (torch.rand(B, 1, hw, hw, generator=g) > 0.3).float()
    l64, s64, g64 = oracle(st, x, tgt, msk, k, fwd, torch.float64)
    l32, s32, g32 = oracle(st, x, tgt, msk, k, fwd, torch.float32)
    print(f"== {name} c{c} k{k} {hw}x{hw} B{B}: oracle fp32 vs fp64: logits {rel(l32,l64):.2e} loss {abs(float(s32)-float(s64))/abs(float(s64)):.2e}")
    for tag, prec, simt in (("ours fp32", "fp32", True), ("ours bf16 simt", "bf16", True), ("ours bf16 tc", "bf16", False)):
        lo, so, go = ours(make, st, x, tgt, msk, k, prec, simt)
        if DETAIL and prec == "fp32":
            for kk in g64:
                if g64[kk] is not None and float(g64[kk].norm()) > 1e-12:
                    print(f"      [detail] {kk:50s} ours {rel(go[kk], g64[kk]):.2e} | oracle32 {rel(g32[kk], g64[kk]):.2e}")
        errs = sorted(((rel(go[kk], g64[kk]), rel(g32[kk], g64[kk]), kk) for kk in g64 if g64[kk] is not None and float(g64[kk].norm()) > 1e-12), reverse=True)
        print(f"  {tag}: logits vs fp64 {rel(lo,l64):.2e}  loss {abs(float(so)-float(s64))/abs(float(s64)):.2e}  worst grads (ours|oracle32 vs fp64):")
        for e, e32, kk in errs[:4]:
            print(f"      {kk:50s} {e:.2e} | {e32:.2e}   |g|={float(g64[kk].norm()):.2e}")
        allg = torch.cat([go[kk].flatten().double().cpu() for kk in g64 if g64[kk] is not None]); allr = torch.cat([g64[kk].flatten() for kk in g64 if g64[kk] is not None])
        print(f"      all-grads-concatenated rel err {float((allg-allr).norm()/allr.norm()):.2e}")
    la, sa, ga = autocast_ref(st, x, tgt, msk, k, fwd)
    allg = torch.cat([ga[kk].flatten().double().cpu() for kk in g64 if g64[kk] is not None]); allr = torch.cat([g64[kk].flatten() for kk in g64 if g64[kk] is not None])
    print(f"  torch autocast-bf16 (cuDNN) same net: logits vs fp64 {rel(la,l64):.2e} loss {abs(float(sa)-float(s64))/abs(float(s64)):.2e} all-grads {float((allg-allr).norm()/allr.norm()):.2e}")

DETAIL = "--detail" in sys.argv
if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    from networks.unet_model import UNet as UA
    from networks.unet import UNet as UB
    if DETAIL:
        case("unet_b", lambda: UB(3, 3), lambda: U.init_unet_b(3, 3, seed=1337), lambda s, x: U.unet_b_forward(s, x, True), 3, 3, 32, 2)
        case("unet_b", lambda: UB(3, 3), lambda: U.init_unet_b(3, 3, seed=1337), lambda s, x: U.unet_b_forward(s, x, True), 3, 3, 64, 2)
        sys.exit(0)
    for hw, B in ((32, 2), (128, 2), (256, 2)):
        case("unet_a", lambda: UA(1, 2), lambda: U.init_unet_a(1, 2, seed=1337), lambda s, x: U.unet_a_forward(s, x, True), 1, 2, hw, B)
    for hw, B in ((32, 2), (128, 2)):
        case("unet_b", lambda: UB(3, 3), lambda: U.init_unet_b(3, 3, seed=1337), lambda s, x: U.unet_b_forward(s, x, True), 3, 3, hw, B)
