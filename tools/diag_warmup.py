"""Per-step wall/GPU time of the first N SSL steps on a fresh box + clocks (why are the first ~15 steps slow?)."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch
import bench
from ustrun import synth as S
from ustrun import engine as E
from ustrun.step import SSLTrainer
t00 = time.time()
model_name, c, k, H, W, Bl, Bu, branch = bench.WORKLOADS[os.environ.get("WORKLOAD", "cfg2")]
student, teacher = bench.make_models(model_name, c, k)
student, teacher = student.cuda().train(), teacher.cuda().train()
tr = SSLTrainer(student, teacher, n_classes=k, branch=branch, base_lr=0.03, max_iterations=60000, threshold=0.95)
tr.iter_num = 30000
host = S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337, branch=branch)
for kk in ("lb_mask", "cut_label", "cut_mask", "box"): host[kk] = host[kk].to(torch.uint8)
host["choice"] = host["choice"].to(torch.int32)
dev = {kk: v.cuda() for kk, v in host.items()}
lq = dev["ulb_w"][:1].contiguous()
if os.environ.get("RESERVE", "1") == "1": print("reserved", E.reserve_pool(fraction=float(os.environ.get("POOL_FRACTION", "0.5")), cap=160 << 30) >> 20, "MiB")
def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:
        return str(e)
print("setup", round(time.time() - t00, 1), "s; smi:", smi(), flush=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.step(dev, lq=lq)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    print(f"step {i:2d}: host enqueue {1e3*(t1-t0):6.1f} ms, gpu {e0.elapsed_time(e1):6.1f} ms, wall {1e3*(t2-t0):6.1f} ms, cudaMalloc retries {st['num_alloc_retries']} segments {st['segment.all.allocated']}, reserved {st['reserved_bytes.all.current']>>20} MiB, peak allocated {st['allocated_bytes.all.peak']>>20} MiB, active {st['active_bytes.all.current']>>20} MiB" + (f"  smi: {smi()}" if i % 5 == 0 else ""), flush=True)
