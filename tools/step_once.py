"""Run a few SSL steps of a bench workload and exit (the target of `ncu` launch lists / `--set full` captures).

    python tools/step_once.py --workload cfg2 --steps 4 [--graph] [--precision bf16]

Uses the same models / synthetic batch as bench.py's measured arm; prints the kernels per step so that the
profiler's -s / -c window can be placed on a steady-state step."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ust-run_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

import bench as B
from ustrun import engine as E
from ustrun import synth as S
from ustrun.step import SSLTrainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(B.WORKLOADS))
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--manifest", default="", help="write the profiled C-ABI calls of the LAST step (class, algorithmic work, shape) as JSON")
    a = ap.parse_args()
    model_name, c, k, H, W, Bl, Bu, branch = B.WORKLOADS[a.workload]
    E.set_precision(a.precision)
    student, teacher = B.make_models(model_name, c, k)
    student, teacher = student.cuda().train(), teacher.cuda().train()
    tr = SSLTrainer(student, teacher, n_classes=k, branch=branch, max_iterations=60000, use_graph=a.graph, lanes=a.lanes)
    tr.iter_num = 30000
    host = S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337, branch=branch)
    dev = {kk: v.cuda() for kk, v in host.items()}
    for kk in ("lb_mask", "cut_label", "cut_mask", "box"):
        dev[kk] = dev[kk].to(torch.uint8)
    dev["choice"] = dev["choice"].to(torch.int32)
    extra = dict(domain_lb=B.DSBN_DOMAINS[0], domain_ulb=B.DSBN_DOMAINS[1]) if model_name.endswith("_dsbn3") else {}
    lq = dev["ulb_w"][:1].contiguous()
    E.reserve_pool(fraction=0.3)
    for i in range(a.steps):
        if a.manifest and i == a.steps - 1:
            E.MANIFEST = []
        out = tr.step({**dev, **extra}, lq=lq)
        torch.cuda.synchronize()
        print(f"step {i}: loss {float(out['loss']):.5f} kernels/step {tr.launches_per_step} (total so far {E.KERNELS})", flush=True)


    if a.manifest:
        import json
        json.dump({"workload": a.workload, "kernels_per_step": tr.launches_per_step, "calls": E.MANIFEST}, open(a.manifest, "w"))


if __name__ == "__main__":
    main()
