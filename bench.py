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
    # name: (model, n_channels, n_classes, H, W, B_l, B_u, branch); model suffix "_dsbn3": DomainSpecificBatchNorm2d over 3 domains
    "cfg1": ("unet_b", 3, 3, 256, 256, 4, 4, "softmax"),                 # BASELINE.json configs[0] (the reference's CPU case)
    "cfg2": ("unet_a", 1, 2, 384, 384, 8, 8, "softmax"),                 # configs[1]: prostate-shaped, the headline
    "cfg2b": ("unet_b", 1, 2, 384, 384, 8, 8, "softmax"),
    "cfg3": ("unet_a_dsbn3", 3, 2, 256, 256, 16, 16, "softmax"),         # configs[2]: BUSI-shaped, DSBN over 3 domains (tensor-bound model)
    "cfg3b": ("unet_b_dsbn3", 3, 2, 256, 256, 16, 16, "softmax"),        # configs[2] on networks/unet.py UNet(norm='dsbn', num_domains=3)
    "cfg3bn": ("unet_a", 3, 2, 256, 256, 16, 16, "softmax"),             # same shapes, plain BatchNorm (round-1 definition of cfg3)
    "cfg4": ("unet_a", 1, 4, 288, 288, 32, 32, "softmax"),               # configs[3]: M&Ms-shaped (train_mnms.py)
    "cfg5": ("unet_a", 3, 2, 512, 512, 64, 64, "softmax"),               # configs[4]
    "fundus8": ("unet_a", 3, 2, 256, 256, 8, 8, "sigmoid"),              # north_star target: fundus 256x256, 8+8, two sigmoid channels (train.py:404-409,649-657)
    "fundus8b": ("unet_b", 3, 2, 256, 256, 8, 8, "sigmoid"),
    "tiny": ("unet_a", 1, 2, 64, 64, 2, 2, "softmax"),
}
METRIC = "ssl_train_step_images_per_sec"
DSBN_DOMAINS = (0, 2)          # (labelled batch's domain, unlabelled batch's domain) of the DSBN workloads


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
    kw = dict(norm="dsbn", num_domains=3) if model.endswith("_dsbn3") else {}
    if model.startswith("unet_a"):
        from networks.unet_model import UNet
        torch.manual_seed(1337)
        student, teacher = UNet(c, k, **kw), UNet(c, k, **kw)
    else:
        from networks.unet import UNet
        torch.manual_seed(1337)
        student, teacher = UNet(c, k, **kw), UNet(c, k, **kw)
    teacher.load_state_dict(student.state_dict())     # teacher = copy of the student after step 0 (alpha=0)
    for p in teacher.parameters():
        p.detach_()
    return student, teacher


def conv_flops_per_image(model, c, k, H, W):
    if model.startswith("unet_a"):
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


def shutdown_dp(trainer):
    """End of a data-parallel run: the captured step graphs hold NCCL kernels of the communicator and have to go BEFORE it
    (destroying the process group with such graphs alive blocks forever); then leave the process without waiting for anything
    else -- the JSON line is out, nothing of value is left to run, and a rank that lingers stalls the launcher."""
    import gc
    import threading

    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    if not dist.is_initialized():
        return
    guard = threading.Timer(30.0, lambda: os._exit(0))      # teardown must never outlive the measurement
    guard.daemon = True
    guard.start()
    torch.cuda.synchronize()
    trainer._graphs.clear()
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    guard.cancel()
    sys.stdout.flush()
    os._exit(0)


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
    elif os.environ.get("USTRUN_BENCH_FORCE_DP"):
        # debugging aid: the complete data-parallel code path (buckets, peer-memory BatchNorm kernels, global loss sums) on ONE GPU
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
        from ustrun.dp import DataParallel
        dp = DataParallel(sync_bn="peer", force=True)
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    dsbn = model_name.endswith("_dsbn3")
    E.set_precision(args.precision)
    student, teacher = make_models(model_name, c, k)
    student, teacher = student.cuda().train(), teacher.cuda().train()
    use_graph = args.graph == "on" or (args.graph == "auto" and (dp is None or dp.graph_safe) and args.workload != "cfg5")
    # auto: the tensor-bound UNet-A gains from 2 lanes (4 thrash the L2), the latency-bound UNet-B from 4
    # data parallel: lanes need the peer-memory BatchNorm path (every cross-rank kernel on one stream)
    dp_lanes_ok = dp is None or getattr(dp, "peer", None) is not None or not dp.sync_bn
    lanes = args.lanes if args.lanes > 0 else (1 if (not dp_lanes_ok or args.workload == "cfg5") else (2 if model_name.startswith("unet_a") else 4))
    trainer = SSLTrainer(student, teacher, n_classes=k, branch=branch, base_lr=0.03, max_iterations=60000, threshold=0.95, dp=dp, use_graph=use_graph,
                         lanes=lanes)
    trainer.iter_num = 30000                      # mid-training: consistency weight 1.0, alpha 0.99
    host = S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337 + rank, branch=branch)
    host["lb_mask"] = host["lb_mask"].to(torch.uint8)
    host["cut_label"] = host["cut_label"].to(torch.uint8)
    host["cut_mask"] = host["cut_mask"].to(torch.uint8)
    host["box"] = host["box"].to(torch.uint8)
    host["choice"] = host["choice"].to(torch.int32)
    pinned = {kk: v.contiguous().pin_memory() for kk, v in host.items()}
    dev = {kk: v.cuda() for kk, v in pinned.items()}
    extra = dict(domain_lb=DSBN_DOMAINS[0], domain_ulb=DSBN_DOMAINS[1]) if dsbn else {}

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
    lq_pinned = pinned["ulb_w"][:1].contiguous().pin_memory()

    def step_resident():
        return trainer.step({**dev, **extra}, lq=lq_dev)

    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values()) + lq_pinned.numel() * 4
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    staged = [trainer.upload({**pinned, **extra}, lq=lq_pinned)]
    loss_ring = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev, loss_seen = [None, None], []

    def step_e2e():
        # the public calls with HOST (pinned) inputs, software-pipelined like an input pipeline with asynchronous logging:
        # step() consumes the batch upload() staged; the NEXT step's inputs are uploaded on the copy stream while this step's
        # kernels run; this step's loss is copied to pinned memory behind its kernels and READ on the host one step later (the
        # host blocks on that copy's event), so the host never idles the GPU.  Per timed step: one upload of h2d_bytes, one loss read.
        i = len(loss_seen) & 1
        out = trainer.step(staged[0])
        staged[0] = trainer.upload({**pinned, **extra}, lq=lq_pinned)
        loss_ring[i].copy_(out["loss"].reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if loss_ev[i ^ 1] is not None:
            loss_ev[i ^ 1].synchronize()                 # the previous step's loss has landed: the user reads it now
            loss_host.copy_(loss_ring[i ^ 1])
        loss_ev[i] = ev
        loss_seen.append(1)
        return out

    dp_parity = dp_parity_record(args, dp, rank, world) if (world > 1 and not args.no_dp_parity) else None
    # one cudaMalloc up front instead of ~40 during the first 20 steps (cfg5 needs > 100 GB of activations + workspaces);
    # graph mode gives the eager pool back before capturing, so it only needs a small reservation
    pool = E.reserve_pool(fraction=(0.6 if args.workload == "cfg5" else 0.5) if not use_graph else 0.1, cap=160 << 30)
    out0 = step_resident()                       # step 0 from the seeded initial weights: kept for the parity record
    parity_loss0 = out0["loss"].detach().clone()
    parity_pl0 = out0["pseudo_label"].detach().clone()
    for _ in range(max(args.warmup, 3 if use_graph else 0) - 1):
        step_resident()
    if use_graph and not trainer.use_graph:      # data parallel: the capture (NCCL calls included) failed and the trainer fell back to eager steps
        use_graph = False
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms = timed(step_resident, args.steps)
    launches = trainer.launches_per_step * args.steps
    if rank == 0:
        sampler.stop()
    # end to end right after the resident loop (same thermal / power state), with its own clock samples
    for _ in range(max(min(3, args.warmup), 3 if use_graph else 1)):
        step_e2e()
    sampler_e2e = ClockSampler(local)
    if rank == 0:
        sampler_e2e.start()
        time.sleep(0.3)
    ms_e2e = timed(step_e2e, args.steps)
    if rank == 0:
        sampler_e2e.stop()
    # per-kernel-class roofline: a separate short EAGER pass with CUDA events around every conv launch (the events
    # serialise the weight-gradient side stream, so this pass is not the one that is timed)
    prof_steps = min(args.steps, 3)
    trainer.use_graph = False
    E.PROFILE_EVENTS = [] if rank == 0 else None
    timed(step_resident, prof_steps)
    prof_events = E.PROFILE_EVENTS
    E.PROFILE_EVENTS = None
    trainer.use_graph = use_graph
    imgs = world * (Bl + Bu)
    value, e2e_value = imgs / (ms / 1e3), imgs / (ms_e2e / 1e3)

    if rank != 0:
        shutdown_dp(trainer)
        return
    pk = peaks()
    f_img = conv_flops_per_image(model_name, c, k, H, W)
    # forwards: 3 teacher + S0 + 3 student branches on B_u, 1 on B_l, 1 low-quality image; backward (dgrad+wgrad = 2x) on 4 branches
    step_flops = (7 * Bu + Bl + 1) * f_img + (Bl + 3 * Bu) * 2 * f_img
    # per-kernel-class roofline from CUDA events recorded around the conv launches in the profiling pass
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
    # DRAM bytes per launch / tensor-pipe activity of the dominant kernel from the committed `ncu --set full` capture (tools/ncu_traffic.py)
    traffic, traffic_note, tensor_pipe = None, "no capture committed", None
    for fname in ("r02_ncu_tc_conv_traffic.json", "r01_ncu_tc_conv_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", fname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_launch"]
            tensor_pipe = tj.get("tensor_pipe_pct")
            traffic_note = (f"mean over {tj['launches']} k_tc_conv launches (cfg2 shapes) of dram__bytes_read+write from {tj['source']}; "
                            f"algorithmic {tj['algorithmic_bytes_per_launch']:.3e} B/launch, ratio {tj['ratio']:.2f}")
            break
    roofline = {"bound": "tensor", "kernel": top, "achieved": kern[top]["tflops"] if top else None, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (kern[top]["tflops"] / peak_tf) if top else None, "traffic": traffic, "traffic_note": traffic_note, "tensor_pipe_pct": tensor_pipe,
                "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                "flops_per_launch": (classes[top][0] / classes[top][2]) if top else None,
                "us_per_launch": (classes[top][1] * 1e3 / classes[top][2]) if top else None,
                "conv_layers": conv_layers, "hbm_kernels": hbm, "hbm_peak_gbs": pk["hbm_gbs"],
                "step_conv_tflops": step_flops / (ms * 1e-3) / 1e12, "step_frac_of_peak": step_flops / (ms * 1e-3) / 1e12 / peak_tf,
                "step_frac_of_burst_peak": step_flops / (ms * 1e-3) / 1e12 / pk["tf_burst"], "kernels": kern}
    cpu = cpu_baseline(args, bounded=True)
    gpu_base, parity = (None, None)
    if world == 1 and not args.no_gpu_baseline:
        gpu_base, parity = gpu_baseline(args, float(parity_loss0), parity_pl0)
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {model_name} {c}x{H}x{W}, {k} classes, {Bl}+{Bu} per GPU, {branch} branch, SSL step",
                       "global_batch": imgs, "parallelism": f"dp{world}", "l2": "working set (GBs of activations per step) >> 126 MB L2; no flush needed", "pool_reserved_gib": round(pool / 2**30, 1),
                       "cuda_graph": bool(use_graph), "graph_error": getattr(trainer, "graph_error", None), "lanes": trainer.lanes, "dsbn_domains": list(DSBN_DOMAINS) if dsbn else None,
                       "sync_bn": (False if (world == 1 or args.no_sync_bn) else ("peer" if dp is not None and dp.peer is not None else "nccl"))},
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "clocks": sampler_e2e.summary(),
                    "how": "trainer.upload(pinned host batch) + trainer.step(), software-pipelined: every timed step uploads one batch (the next step's, on a copy stream, overlapping this step's kernels) and reads one loss on the host (the previous step's: the host blocks on that D2H copy's event, not on the whole stream)"},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "parity": parity,
            "dp_parity": dp_parity}
    print(json.dumps(line))
    shutdown_dp(trainer)


def dp_parity_record(args, dp, rank, world):
    """Pre-timing data-parallel parity step (N > 1): `world` ranks x (2+2) images through the DataParallel step (NCCL gradient
    all-reduce, cross-rank BatchNorm statistics, global CE/Dice sums) against ONE process stepping on the concatenated batch
    (rank 0, hooks detached) from the same seeded weights: step-0 loss and the updated student weights.  Both sides are this
    repo's sm_100a step; what is checked is that G ranks x B images == one device with G*B images."""
    import torch
    import torch.distributed as dist
    from ustrun import bridge
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    model_name, c, k, H, W, _, _, branch = WORKLOADS[args.workload]
    B = 2
    dsbn = model_name.endswith("_dsbn3")
    extra = dict(domain_lb=DSBN_DOMAINS[0], domain_ulb=DSBN_DOMAINS[1]) if dsbn else {}
    full = S.synthetic_batch(c, k, H, W, B * world, B * world, seed=4242, branch=branch)
    sl = slice(rank * B, (rank + 1) * B)
    local_b = {kk: (v[sl] if kk in ("lb_x", "lb_mask", "ulb_w", "ulb_s", "move_transx", "box", "choice") else v) for kk, v in full.items()}
    s_dp, t_dp = make_models(model_name, c, k)
    tr = SSLTrainer(s_dp.cuda().train(), t_dp.cuda().train(), n_classes=k, branch=branch, max_iterations=60000, threshold=0.6, dp=dp)
    tr.iter_num = 30000
    out = tr.step({**{kk: v.cuda() for kk, v in local_b.items()}, **extra})
    loss_dp = float(out["loss"])
    torch.cuda.synchronize()
    rec = None
    if rank == 0:
        saved = (bridge.BN_SYNC, bridge.BN_WORLD)
        bridge.BN_SYNC, bridge.BN_WORLD = None, 1
        try:
            s_1, t_1 = make_models(model_name, c, k)
            tr1 = SSLTrainer(s_1.cuda().train(), t_1.cuda().train(), n_classes=k, branch=branch, max_iterations=60000, threshold=0.6)
            tr1.iter_num = 30000
            out1 = tr1.step({**{kk: v.cuda() for kk, v in full.items()}, **extra})
            loss_1 = float(out1["loss"])
            cat = lambda m: torch.cat([q.detach().double().flatten() for q in m.parameters()])
            w_dp, w_1 = cat(tr.model), cat(tr1.model)
            rec = {"what": f"{world} ranks x (2+2) images (data-parallel step) vs one process on the concatenated {2 * B * world}-image batch, same seeded weights, {args.precision}",
                   "loss_dp": loss_dp, "loss_single": loss_1, "loss_rel_err": abs(loss_dp - loss_1) / abs(loss_1),
                   "student_weights_rel_diff_after_step": float((w_dp - w_1).norm() / w_1.norm()),
                   "masks_identical": bool(torch.equal(out["mask_w"], out1["mask_w"][sl]))}
            del s_1, t_1, tr1, out1
        finally:
            bridge.BN_SYNC, bridge.BN_WORLD = saved
    dist.barrier()
    del tr, s_dp, t_dp, out
    torch.cuda.empty_cache()
    return rec


def oracle_models(model_name, c, k, device="cpu"):
    """State dicts + forward of the oracle port for a workload (same seed => same initial weights as make_models)."""
    from oracle import unet_ref as U
    dsbn = model_name.endswith("_dsbn3")
    kw = dict(norm="dsbn", num_domains=3) if dsbn else {}
    if model_name.startswith("unet_a"):
        st = U.init_unet_a(c, k, seed=1337, **kw)
        fwd = (lambda s, x, dl: U.unet_a_forward(s, x, True, domain_label=dl)) if dsbn else (lambda s, x: U.unet_a_forward(s, x, True))
    else:
        st = U.init_unet_b(c, k, seed=1337, **kw)
        fwd = (lambda s, x, dl: U.unet_b_forward(s, x, True, domain_label=dl)) if dsbn else (lambda s, x: U.unet_b_forward(s, x, True))
    st = {n: v.to(device) for n, v in st.items()}
    return st, {n: v.clone() for n, v in st.items()}, fwd, dsbn


def oracle_domains(dsbn, B_l, B_u):
    if not dsbn:
        return None
    import torch
    u, l = torch.full((B_u,), DSBN_DOMAINS[1], dtype=torch.long), torch.full((B_l,), DSBN_DOMAINS[0], dtype=torch.long)
    return dict(t1=u, t2=u, t3=l, s0=u, lb=l, ul=u, lu=l, s=u, lq=u)


def cpu_step_time(model_name, c, k, H, W, Bl, Bu, steps, warmup, branch="softmax"):
    """Time the oracle port (torch CPU fp32 restatement of the reference step) on the host cores."""
    import torch
    from oracle import ssl_step_ref as S
    torch.set_num_threads(os.cpu_count() or 1)
    st_s, st_t, fwd, dsbn = oracle_models(model_name, c, k)
    batch = S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337, branch=branch)
    bufs, times = {}, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        S.ssl_step(fwd, st_s, st_t, bufs, batch, n_classes=k, branch=branch, iter_num=30000 + i, max_iterations=60000, lq=batch["ulb_w"][:1],
                   domains=oracle_domains(dsbn, Bl, Bu))
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def gpu_baseline(args, our_loss0, our_pl0):
    """Baseline leg (N=1 only, like cpu_baseline): the reference's step as PyTorch eager + cuDNN executes it on THIS GPU -- the
    oracle port moved to CUDA -- in fp32 (TF32 off), bf16 autocast and fp16 autocast (what the reference ships, train.py:30,551).
    Also yields the parity record: our bf16 step-0 loss / pseudo labels against the fp32 eager step from the same seeded weights
    and inputs.  The oracle is the yardstick here, never the thing measured as 'ours'."""
    import torch
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    if (Bl + Bu) * H * W > 16 * 384 * 384 * 2:
        return {"skipped": "eager autograd of this workload does not fit next to the measured arm's pool (SURVEY H3)"}, None
    try:
        from oracle import ssl_step_ref as S
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        torch.cuda.empty_cache()
        batch = {kk: v.cuda() for kk, v in S.synthetic_batch(c, k, H, W, Bl, Bu, seed=1337, branch=branch).items()}
        res, parity = {}, None
        for name, dt, scale in (("fp32", None, 1.0), ("bf16_autocast", torch.bfloat16, 1.0), ("fp16_autocast", torch.float16, 1024.0)):
            st_s, st_t, fwd, dsbn = oracle_models(model_name, c, k, "cuda")
            bufs, ts = {}, []
            for i in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.autocast("cuda", dtype=dt, enabled=dt is not None):
                    out = S.ssl_step(fwd, st_s, st_t, bufs, batch, n_classes=k, branch=branch, iter_num=30000 + i, max_iterations=60000, lq=batch["ulb_w"][:1],
                                     domains=oracle_domains(dsbn, Bl, Bu), loss_scale=scale)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
                if i == 0 and name == "fp32":
                    ref = float(out["loss"])
                    agree = float((our_pl0.reshape(-1).long() == out["pseudo_label"].reshape(-1).long()).float().mean())
                    parity = {"what": "step 0 from the seeded initial weights: ours (" + args.precision + ") vs the oracle port in fp32 on this GPU (cuDNN, TF32 off)",
                              "loss_ours": our_loss0, "loss_fp32_eager": ref, "loss_rel_err": abs(our_loss0 - ref) / abs(ref), "pseudo_label_agreement": agree,
                              "full_size_tests": "tests/test_parity_fullsize_gpu.py (warmed weights: logits, gradients, state after the step, 50-step trajectory)"}
            res[name] = {"ms_per_step": min(ts[1:]) * 1e3, "images_per_s": (Bl + Bu) / min(ts[1:])}
            del st_s, st_t, bufs, out
            torch.cuda.empty_cache()
        res["what"] = "oracle port of the reference step under torch eager + cuDNN on this GPU, best of 2 after 1 warm-up step"
        return res, parity
    except Exception as e:
        return {"failed": f"{type(e).__name__}: {e}"}, None


def cpu_baseline(args, bounded=True):
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    B = 1 if bounded else Bl
    try:
        t = cpu_step_time(model_name, c, k, H, W, B, B, steps=1, warmup=0, branch=branch)
        return {"value": 2 * B / t, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"1 step of the oracle port at B_l=B_u={B} (+ the batch-1 low-quality forward; same {c}x{H}x{W} shape, fp32, {os.cpu_count()} torch threads); "
                          f"{t:.1f} s; bounded sample of the {Bl}+{Bu} step, see --impl reference for the full-size CPU step"}
    except Exception as e:          # never lose the GPU line because the CPU leg failed
        return {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference (pure
    Python, no build) does not travel to the GPU box, so this times the oracle port, which is pinned
    bit-for-bit against the reference modules (tests/test_oracle_golden.py).  Same workload as the measured
    arm (full B_l + B_u step incl. the batch-1 low-quality forward) whenever 1 warm-up + 2 timed steps fit the
    time budget (--ref-budget seconds, estimated from a 1+1 probe step); otherwise the largest batch that fits."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, c, k, H, W, Bl, Bu, branch = WORKLOADS[args.workload]
    t_probe = cpu_step_time(model_name, c, k, H, W, 1, 1, steps=1, warmup=0, branch=branch)
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    per_image = t_probe / 2.0                              # optimistic: larger batches thread better
    fit = int(args.ref_budget / max(per_image * (steps + warmup), 1e-9)) // 2
    B_l, B_u = (Bl, Bu) if fit >= max(Bl, Bu) else (max(1, min(Bl, fit)), max(1, min(Bu, fit)))
    same = (B_l, B_u) == (Bl, Bu)
    t = cpu_step_time(model_name, c, k, H, W, B_l, B_u, steps=steps, warmup=warmup, branch=branch)
    value = (B_l + B_u) / t
    world = int(os.environ.get("WORLD_SIZE", "1"))
    desc = f"{Bl}+{Bu} per step" if same else f"bounded sample {B_l}+{B_u} of the {Bl}+{Bu} step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {model_name} {c}x{H}x{W}, {k} classes, {desc}, {branch} branch, SSL step on CPU", "same_batch_as_gpu_arm": same,
                       "global_batch": B_l + B_u},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{steps} timed step(s) after {warmup} warm-up at B_l={B_l}, B_u={B_u}, fp32, {os.cpu_count()} torch threads (probe step at 1+1: {t_probe:.1f} s)"},
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
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-cuDNN baseline / parity leg (N=1)")
    ap.add_argument("--no-dp-parity", action="store_true", help="skip the pre-timing data-parallel parity step (N>1)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="replay the step as one CUDA graph (auto: on, except cfg5 and data parallel with per-layer NCCL BatchNorm statistics)")
    ap.add_argument("--lanes", type=int, default=0, help="CUDA streams the independent forwards / loss branches of a step are spread over (0: auto = 2 for UNet-A, 4 for UNet-B, 1 for cfg5)")
    ap.add_argument("--ref-budget", type=float, default=420.0, help="--impl reference: seconds of CPU time the whole run may take")
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
