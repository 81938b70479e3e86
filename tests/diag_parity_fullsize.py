"""Full-size parity diagnostics (GPU).  Not part of the product; writes gpurun_out/parity_fullsize.json.

    python tests/diag_parity_fullsize.py [case ...] [--warm N] [--traj N]

For every case: warm the network with the sm_100a step, then step ONCE from that state with
  float64 oracle (truth) | fp32 oracle | bf16-autocast oracle | fp16-autocast oracle (+ static loss scale) | ours bf16 | ours fp32
and print every implementation's error against float64: loss terms, logits of the eight forwards, label/mask planes,
gradients (all / median / worst tensor), parameter update, running statistics.  --traj N: N further steps of ours-bf16 and the
fp32 oracle from the warmed state on identical batches (loss curves)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ust-run_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

import parity_lib as P
from synth_tasks import to_device_batch

CASES = {
    "cfg2": P.Case("cfg2", "unet_a", 1, 2, 384, 384, 8),
    "fundus8": P.Case("fundus8", "unet_a", 3, 2, 256, 256, 8, branch="sigmoid"),
    "cfg1": P.Case("cfg1", "unet_b", 3, 3, 256, 256, 4),
    "cfg3b": P.Case("cfg3b", "unet_b_dsbn3", 3, 2, 256, 256, 16),
    "cfg3": P.Case("cfg3", "unet_a_dsbn3", 3, 2, 256, 256, 16),
    "cfg4": P.Case("cfg4", "unet_a", 1, 4, 288, 288, 32),
    "small_a": P.Case("small_a", "unet_a", 1, 2, 64, 64, 2),
    "small_b": P.Case("small_b", "unet_b_dsbn3", 3, 2, 64, 64, 2),
}


def log(*a):
    print(*a, flush=True)


def summarize(tag, r):
    log(f"  {tag:18s} loss {r['loss']:.2e} terms {r['sup_loss']:.1e}/{r['unsup_loss_ul']:.1e}/{r['unsup_loss_lu']:.1e}/{r['unsup_loss_s']:.1e}  "
        f"logits max {r['logits_max']:.2e} (t1 {r['logits']['t1']:.1e} lb {r['logits']['lb']:.1e} s {r['logits']['s']:.1e})  planes>= {r['planes_min_agreement']:.4f}  "
        f"grads all {r['grads_all']:.2e} med {r['grads_median']:.2e} worst {r['grads_worst'][0][0]:.2e} ({r['grads_worst'][0][1]})  "
        f"update {r['update']:.2e}  running {r['running_stats_max']:.1e}")


def run_case(case, warm_steps, traj, results, f64=True):
    t0 = time.time()
    log(f"== {case.name}: {case.kind} {case.c}x{case.H}x{case.W} k={case.k} B={case.B}+{case.B} {case.branch}")
    student, teacher, tr = P.warm(case, warm_steps, log=log)
    st_s, st_t, bufs = P.export_state(student), P.export_state(teacher), P.export_momentum(tr)
    it, lr = tr.iter_num, tr.lr
    del student, teacher, tr
    torch.cuda.empty_cache()
    batch = case.batch(777)
    res = {"warm_steps": warm_steps}
    truth_dtype = torch.float64 if f64 else torch.float32
    ref, ref_s, ref_t = P.oracle_step(case, st_s, st_t, bufs, batch, it, lr, dtype=truth_dtype)
    res["truth"] = "float64" if f64 else "float32"
    res["mask_mean"] = {p: float(ref[p].float().mean()) for p in ("mask", "mask_w", "mask_ul", "mask_lu")}
    res["loss_truth"] = float(ref["loss"])
    log(f"  truth ({res['truth']}): loss {float(ref['loss']):.5f} mask means {res['mask_mean']}  [{time.time()-t0:.0f}s]")
    impls = []
    if f64:
        impls.append(("oracle_fp32", dict(dtype=torch.float32)))
    impls += [("oracle_bf16_autocast", dict(dtype=torch.float32, autocast=torch.bfloat16)),
              ("oracle_fp16_autocast", dict(dtype=torch.float32, autocast=torch.float16, loss_scale=65536.0))]
    for tag, kw in impls:
        try:
            o, s_after, _ = P.oracle_step(case, st_s, st_t, bufs, batch, it, lr, **kw)
            if not all(torch.isfinite(g).all() for g in o["grads"].values() if g is not None):
                log(f"  {tag}: non-finite gradients at loss scale {kw.get('loss_scale')}; retrying with 1024")
                kw = dict(kw, loss_scale=1024.0)
                o, s_after, _ = P.oracle_step(case, st_s, st_t, bufs, batch, it, lr, **kw)
            res[tag] = P.compare_step(ref, ref_s, st_s, *P.oracle_as_got(o, s_after))
            summarize(tag, res[tag])
            del o, s_after
        except Exception as e:
            log(f"  {tag}: FAILED {type(e).__name__}: {e}")
        torch.cuda.empty_cache()
    for tag, prec in (("ours_bf16", "bf16"), ("ours_fp32", "fp32")):
        try:
            o, g, s_after, _ = P.ours_step(case, st_s, st_t, bufs, batch, it, lr, prec)
            res[tag] = P.compare_step(ref, ref_s, st_s, o, g, s_after)
            summarize(tag, res[tag])
            del o, g, s_after
        except Exception as e:
            import traceback
            traceback.print_exc()
            log(f"  {tag}: FAILED {type(e).__name__}: {e}")
        torch.cuda.empty_cache()
    del ref, ref_s, ref_t
    torch.cuda.empty_cache()
    if traj:
        from oracle import ssl_step_ref as S
        from ustrun import engine as E
        E.set_precision("bf16")
        student, teacher = P.make_pair(case.kind, case.c, case.k)
        student.load_state_dict(st_s), teacher.load_state_dict(st_t)
        tr = case.trainer(student, teacher)
        tr.iter_num, tr.lr = it, lr
        names = [n for n, _ in student.named_parameters()]
        for i, n in enumerate(names):
            if bufs.get(n) is not None:
                p = tr.params[i]
                tr.opt.flat_buf[tr.opt.offsets[i]: tr.opt.offsets[i] + p.numel()].view(p.shape).copy_(bufs[n])
                tr.opt.first[i] = False
        s32 = {n: v.clone().cuda() for n, v in st_s.items()}
        t32 = {n: v.clone().cuda() for n, v in st_t.items()}
        b32 = {n: (None if v is None else v.clone().cuda()) for n, v in bufs.items()}
        dom = P.domains_for(case.kind, case.d_lb, case.d_ulb, case.B, case.B)
        lo, lr_, lr_ref = [], lr, lr
        for j in range(traj):
            bt = case.batch(5000 + j)
            o = tr.step({**to_device_batch(bt), **case.extra()})
            r = S.ssl_step(P.oracle_forward(case.kind), s32, t32, b32, {n: v.cuda() for n, v in bt.items()}, n_classes=case.k, branch=case.branch,
                           iter_num=it + j, max_iterations=case.max_iterations, lr=lr_ref, threshold=case.threshold, domains=dom)
            lr_ref = r["next_lr"]
            lo.append((float(o["loss"]), float(r["loss"])))
        res["trajectory"] = lo
        errs = [abs(a - b) / abs(b) for a, b in lo]
        log(f"  trajectory {traj} steps: loss rel err max {max(errs):.2e} mean {sum(errs)/len(errs):.2e} last {errs[-1]:.2e}; ours {lo[0][0]:.4f}->{lo[-1][0]:.4f} ref {lo[0][1]:.4f}->{lo[-1][1]:.4f}")
        wrel = P.rel(torch.cat([p.detach().flatten() for p in student.parameters()]), torch.cat([s32[n].flatten() for n in names]))
        res["trajectory_weights_rel"] = wrel
        log(f"  weights after the trajectory: rel diff {wrel:.2e}")
        del student, teacher, tr, s32, t32, b32
        torch.cuda.empty_cache()
    res["seconds"] = time.time() - t0
    results[case.name] = res


def main():
    args = sys.argv[1:]
    warm_steps, traj, names = 60, 0, []
    i = 0
    while i < len(args):
        if args[i] == "--warm":
            warm_steps = int(args[i + 1]); i += 2
        elif args[i] == "--traj":
            traj = int(args[i + 1]); i += 2
        else:
            names.append(args[i]); i += 1
    names = names or ["cfg2"]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    results = {}
    out = os.path.join(ROOT, "gpurun_out", os.environ.get("USTRUN_PARITY_OUT", "parity_fullsize.json"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    for n in names:
        try:
            run_case(CASES[n], warm_steps, traj, results, f64=(n != "cfg4"))
        except Exception as e:
            import traceback
            traceback.print_exc()
            results[n] = {"failed": f"{type(e).__name__}: {e}"}
        json.dump(results, open(out, "w"), indent=1)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
