"""cProfile of the host side of SSLTrainer.step (where do the ~17 us per launch go?)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch
import bench
from ustrun import synth as S
from ustrun import engine as E
from ustrun.step import SSLTrainer
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
E.reserve_pool(24 << 30)
for _ in range(4): tr.step(dev, lq=lq)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3): tr.step(dev, lq=lq)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
