#!/usr/bin/env python
"""bench.py -- SSL train-step throughput (images/sec) of the B200-native UST-RUN hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = everything in train.py:638-856 on one batch of synthetic input (3 teacher forwards,
5 student forwards, fused pseudo labels, 4 CE+Dice terms, backward, SGD, EMA [, gradient
all-reduce]).  images/sec = G * (B_l + B_u) / t_step.  Prints ONE JSON line on rank 0.

Workload (config.workload): BASELINE.json configs[1] -- prostate-shaped 1x384x384, 2 classes,
8 labelled + 8 unlabelled per GPU, bf16, UNet-A (networks/unet_model.UNet, the model train.py builds).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "ust-run_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (model, n_channels, n_classes, H, W, B_l, B_u, branch)
    "cfg1": ("unet_b", 3, 3, 256, 256, 4, 4, "softmax"),
    "cfg2": ("unet_a", 1, 2, 384, 384, 8, 8, "softmax"),
    "cfg2b": ("unet_b", 1, 2, 384, 384, 8, 8, "softmax"),
    "cfg3": ("unet_a", 3, 2, 256, 256, 16, 16, "softmax"),
    "cfg4": ("unet_a", 1, 4, 288, 288, 32, 32, "softmax"),
    "cfg5": ("unet_a", 3, 2, 512, 512, 64, 64, "softmax"),
    "tiny": ("unet_a", 1, 2, 64, 64, 2, 2, "softmax"),
}
METRIC = "ssl_train_step_images_per_sec"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.kill()       # exact PID we started
            except Exception:
                pass

    def summary(self):
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_models(model, c, k):
    import torch
    if model == "unet_a":
        from networks.unet_model import UNet
        torch.manual_seed(1337)
        student, teacher = UNet(c, k), UNet(c, k)
    else:
        from networks.unet import UNet
        torch.manual_seed(1337)
        student, teacher = UNet(c, k), UNet(c, k)
    teacher.load_state_dict(student.state_dict())     # teacher = copy of the student after step 0 (alpha=0)
    for p in teacher.parameters():
        p.detach_()
    return student, teacher


def conv_flops_per_image(model, c, k, H, W):
    if model == "unet_a":
        from ustrun.synth import conv_flops_unet_a
        return conv_flops_unet_a(c, k, H, W)
    # UNet-B: trace from the layer table (3 convs per ConvD, conv1/conv2(1x1)/conv3 per ConvU, 3x3 head)
    n, f, h, w, cin = 16, 0, H, W, c
    for i in range(5):
        co = n << i
        if i:
            h, w = h // 2, w // 2
        f += 2 * 9 * h * w * (cin * co + 2 * co * co)
        cin = co
    planes = 16 * n
    for i in range(4):
        if i:
            f += 2 * 9 * h * w * (2 * planes) * planes
        h, w = h * 2, w * 2
        f += 2 * h * w * planes * (planes // 2)
        f += 2 * 9 * h * w * planes * planes
        planes //= 2
    f += 2 * 9 * H * W * (2 * n) * k
    return f


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ustrun import synth as S               # synthetic input generator (product side; no oracle on this arm)
    from ustrun import engine as E
    from ustrun.step import SSLTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dp = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from ustrun.dp import DataParallel
        dp = DataParallel(sync_bn=False if args.no_sync_bn else (args.sync_bn if args.sync_bn != "auto" else True))
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    E.set_precision(args.precision)
    student, teacher = make_models(model_name, c, k)
    student, teacher = student.cuda().train(), teacher.cuda().train()
    trainer = SSLTrainer(student, teacher, n_classes=k, branch=branch, base_lr=0.03, max_iterations=60000, threshold=0.95, dp=dp)
    trainer.iter_num = 30000                      # mid-training: consistency weight 1.0, alpha 0.99
    host = S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337 + rank, branch=branch)
    host["lb_mask"] = host["lb_mask"].to(torch.uint8)
    host["cut_label"] = host["cut_label"].to(torch.uint8)
    host["cut_mask"] = host["cut_mask"].to(torch.uint8)
    host["box"] = host["box"].to(torch.uint8)
    host["choice"] = host["choice"].to(torch.int32)
    pinned = {kk: v.contiguous().pin_memory() for kk, v in host.items()}
    dev = {kk: v.cuda() for kk, v in pinned.items()}

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms) / steps

    lq_dev = dev["ulb_w"][:1].contiguous()       # steady state: the batch-1 low-quality forward (train.py:740)

    def step_resident():
        trainer.step(dev, lq=lq_dev)

    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def step_e2e():
        d = {kk: v.cuda(non_blocking=True) for kk, v in pinned.items()}
        out = trainer.step(d, lq=d["ulb_w"][:1])
        loss_host.copy_(out["loss"].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the user reads the loss every step

    # one cudaMalloc up front instead of ~40 during the first 20 steps (cfg5 needs > 100 GB of activations + workspaces)
    pool = E.reserve_pool(fraction=0.6 if args.workload == "cfg5" else 0.5, cap=160 << 30)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = E.KERNELS
    ms = timed(step_resident, args.steps)
    launches = E.KERNELS - l0
    if rank == 0:
        sampler.stop()
    # per-kernel-class roofline: a separate short pass with CUDA events around every conv launch (the events
    # serialise the weight-gradient side stream, so this pass is not the one that is timed)
    prof_steps = min(args.steps, 3)
    E.PROFILE_EVENTS = [] if rank == 0 else None
    timed(step_resident, prof_steps)
    prof_events = E.PROFILE_EVENTS
    E.PROFILE_EVENTS = None
    for _ in range(min(2, args.warmup)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    imgs = world * (Bl + Bu)
    value, e2e_value = imgs / (ms / 1e3), imgs / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    f_img = conv_flops_per_image(model_name, c, k, H, W)
    # forwards: 3 teacher + S0 + 3 student branches on B_u, 1 on B_l, 1 low-quality image; backward (dgrad+wgrad = 2x) on 4 branches
    step_flops = (7 * Bu + Bl + 1) * f_img + (Bl + 3 * Bu) * 2 * f_img
    # per-kernel-class roofline from CUDA events recorded around the conv launches in the timed region
    classes = {}
    torch.cuda.synchronize()
    for name, flops, ev0, ev1 in (prof_events or []):
        cl = classes.setdefault(name, [0.0, 0.0, 0])
        cl[0] += flops
        cl[1] += ev0.elapsed_time(ev1)
        cl[2] += 1
    kern = {n: {"tflops": (v[0] / (v[1] * 1e-3) / 1e12) if v[1] > 0 else None, "ms_per_step": v[1] / prof_steps, "launches_per_step": v[2] / prof_steps}
            for n, v in classes.items() if not n.startswith("hbm_")}
    # HBM-bound kernel classes: achieved algorithmic GB/s against the measured copy bandwidth
    hbm = {n[4:]: {"gbs": v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else None, "frac_of_hbm_peak": v[0] / (v[1] * 1e-3) / 1e9 / pk["hbm_gbs"] if v[1] > 0 else None,
                   "ms_per_step": v[1] / prof_steps, "launches_per_step": v[2] / prof_steps}
           for n, v in classes.items() if n.startswith("hbm_")}
    top = max(kern, key=lambda n: kern[n]["ms_per_step"]) if kern else None
    peak_tf = pk["tf_sustained"]
    # all tensor-core conv launches of a step together (forward, dgrad, wgrad, transposed convs)
    tc = [v for n, v in classes.items() if n.startswith("tc_")]
    tc_flops, tc_ms = sum(v[0] for v in tc), sum(v[1] for v in tc)
    conv_layers = {"tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None, "ms_per_step": tc_ms / prof_steps,
                   "frac_of_sustained_peak": tc_flops / (tc_ms * 1e-3) / 1e12 / pk["tf_sustained"] if tc_ms > 0 else None,
                   "frac_of_burst_peak": tc_flops / (tc_ms * 1e-3) / 1e12 / pk["tf_burst"] if tc_ms > 0 else None}
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (tools/ncu_traffic.py)
    traffic, traffic_note = None, "no capture committed"
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_tc_conv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_launch"]
        traffic_note = (f"mean over {tj['launches']} k_tc_conv launches (13 layers x fwd/dgrad, cfg2 shapes) of dram__bytes_read+write from {tj['source']}; "
                        f"algorithmic {tj['algorithmic_bytes_per_launch']:.3e} B/launch, ratio {tj['ratio']:.2f}")
    roofline = {"bound": "tensor", "kernel": top, "achieved": kern[top]["tflops"] if top else None, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (kern[top]["tflops"] / peak_tf) if top else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                "flops_per_launch": (classes[top][0] / classes[top][2]) if top else None,
                "us_per_launch": (classes[top][1] * 1e3 / classes[top][2]) if top else None,
                "conv_layers": conv_layers, "hbm_kernels": hbm, "hbm_peak_gbs": pk["hbm_gbs"],
                "step_conv_tflops": step_flops / (ms * 1e-3) / 1e12, "step_frac_of_peak": step_flops / (ms * 1e-3) / 1e12 / peak_tf,
                "kernels": kern}
    cpu = cpu_baseline(args, bounded=True)
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {model_name} {c}x{H}x{W}, {k} classes, {Bl}+{Bu} per GPU, {branch} branch, SSL step",
                       "global_batch": imgs, "parallelism": f"dp{world}", "l2": "working set (GBs of activations per step) >> 126 MB L2; no flush needed", "pool_reserved_gib": round(pool / 2**30, 1),
                       "sync_bn": (False if (world == 1 or args.no_sync_bn) else ("peer" if dp is not None and dp.peer is not None else "nccl"))},
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_step_time(model_name, c, k, H, W, B, steps, warmup, branch="softmax"):
    """Time the oracle port (torch CPU fp32 restatement of the reference step) on the host cores."""
    import torch
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    torch.set_num_threads(os.cpu_count() or 1)
    if model_name == "unet_a":
        st_s, st_t = U.init_unet_a(c, k, seed=1337), U.init_unet_a(c, k, seed=1337)
        fwd = lambda s, x: U.unet_a_forward(s, x, True)
    else:
        st_s, st_t = U.init_unet_b(c, k, seed=1337), U.init_unet_b(c, k, seed=1337)
        fwd = lambda s, x: U.unet_b_forward(s, x, True)
    batch = S.synthetic_batch(c, k, H, W, B, B, seed=1337, branch=branch)
    bufs, times = {}, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        S.ssl_step(fwd, st_s, st_t, bufs, batch, n_classes=k, branch=branch, iter_num=30000 + i, max_iterations=60000)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_baseline(args, bounded=True):
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    B = 1 if bounded else Bl
    try:
        t = cpu_step_time(model_name, c, k, H, W, B, steps=1, warmup=0, branch=branch)
        return {"value": 2 * B / t, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"1 untimed-warmup-free step of the oracle port at B_l=B_u={B} (same {c}x{H}x{W} shape, fp32, {os.cpu_count()} torch threads); "
                          f"{t:.1f} s; per-image work is batch-independent"}
    except Exception as e:          # never lose the GPU line because the CPU leg failed
        return {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference (pure
    Python, no build) does not travel to the GPU box, so this times the oracle port, which is pinned
    bit-for-bit against the reference modules (tests/test_oracle_golden.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    B = 1
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    t = cpu_step_time(model_name, c, k, H, W, B, steps=steps, warmup=warmup, branch=branch)
    value = 2 * B / t
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {model_name} {c}x{H}x{W}, {k} classes, bounded sample {B}+{B} of the {Bl}+{Bu} step, {branch} branch, SSL step on CPU"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{steps} step(s) at B_l=B_u={B}, fp32, {os.cpu_count()} torch threads"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-sync-bn", action="store_true")
    ap.add_argument("--sync-bn", default="auto", choices=["auto", "peer", "nccl"], help="cross-rank BN statistics: fused peer-memory kernel or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.no_cpu_baseline:
        global cpu_baseline
        cpu_baseline = lambda a, bounded=True: None
    run_ours(args)


if __name__ == "__main__":
    main()
