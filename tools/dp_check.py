"""Data-parallel parity on real GPUs (run under torchrun, world_size W):
W ranks x (B_l+B_u) images with NCCL gradient all-reduce, cross-rank BN statistics and global
CE/Dice sums must reproduce ONE process stepping on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ust-run_b200"))
import torch, torch.distributed as dist
from ustrun import synth as S
from ustrun import engine as E
from ustrun.step import SSLTrainer
from ustrun.dp import DataParallel
from networks.unet_model import UNet
from networks.unet import UNet as UNetB

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    precision = os.environ.get("USTRUN_PRECISION", "fp32")
    E.set_precision(precision)
    # USTRUN_DP_MODEL=b_dsbn: networks/unet.py UNet(norm='dsbn', num_domains=3) with per-forward domain labels -- the cross-rank
    # statistics then belong to the selected domain's BatchNorm on every rank (BASELINE.json configs[2])
    dsbn = os.environ.get("USTRUN_DP_MODEL", "a") == "b_dsbn"
    c, k, hw, B = (3, 2, 64, 2) if dsbn else (1, 2, 64, 2)
    extra = dict(domain_lb=0, domain_ulb=2) if dsbn else {}
    full = S.synthetic_batch(c, k, hw, hw, B * world, B * world, seed=1337)
    full["choice"] = full["choice"] % (B * world)
    def models():
        torch.manual_seed(1337)
        if dsbn:
            s, t = UNetB(c, k, norm="dsbn", num_domains=3), UNetB(c, k, norm="dsbn", num_domains=3)
        else:
            s, t = UNet(c, k), UNet(c, k)
        t.load_state_dict(s.state_dict())
        for p in t.parameters(): p.detach_()
        return s.cuda().train(), t.cuda().train()
    # --- data parallel: rank r owns images [r*B, (r+1)*B) of every batch tensor; the CutMix partner pool (cut_*) is replicated
    sl = slice(rank * B, (rank + 1) * B)
    local_batch = {kk: (v[sl] if kk in ("lb_x", "lb_mask", "ulb_w", "ulb_s", "move_transx", "box", "choice") else v) for kk, v in full.items()}
    dp = DataParallel(sync_bn=os.environ.get("USTRUN_SYNC_BN", "peer"), global_loss=True, bucket_bytes=8 << 20)
    s_dp, t_dp = models()
    lanes = int(os.environ.get("USTRUN_DP_LANES", "1"))
    graph = os.environ.get("USTRUN_DP_USE_GRAPH", "0") == "1"
    tr = SSLTrainer(s_dp, t_dp, n_classes=k, threshold=0.6, dp=dp, lanes=lanes, use_graph=graph)
    tr.iter_num = 3000
    outs = [tr.step({**{kk: v.cuda() for kk, v in local_batch.items()}, **extra}) for _ in range(2)]
    torch.cuda.synchronize()
    if dp.peer is not None:
        dp.peer.check()
    if lanes > 1 or graph:
        # multi-lane (and, with USTRUN_DP_USE_GRAPH=1, CUDA-graph replayed: steps 3-5) data-parallel step == single-lane eager
        # data-parallel step, bit for bit, on fresh model pairs
        finals = []
        for ln, gr in ((lanes, graph), (1, False)):
            s_x, t_x = models()
            tr_x = SSLTrainer(s_x, t_x, n_classes=k, threshold=0.6, dp=dp, lanes=ln, use_graph=gr)
            tr_x.iter_num = 3000
            for _ in range(5):
                tr_x.step({**{kk: v.cuda() for kk, v in local_batch.items()}, **extra})
            torch.cuda.synchronize()
            dp.peer.check()
            if gr:
                assert tr_x.use_graph and len(tr_x._graphs) == 1, getattr(tr_x, "graph_error", None)
            finals.append([q.detach().clone() for q in list(s_x.parameters()) + list(s_x.buffers()) + list(t_x.parameters()) + list(t_x.buffers())])
        bit = all(torch.equal(a, b) for a, b in zip(*finals))
        flags = [torch.zeros(1, device="cuda") for _ in range(world)]
        dist.all_gather(flags, torch.tensor([float(bit)], device="cuda"))
        if rank == 0:
            print(f"dp_check lanes={lanes} graph={graph}: == single-lane eager data-parallel step after 5 steps, bit-identical on every rank: {[bool(f.item()) for f in flags]}")
        assert bit, "multi-lane / graph-replayed data-parallel step differs from the single-lane eager one"
    if rank == 0:
        print("BN statistics path:", "peer-memory fused finalize" if dp.peer is not None else "NCCL all-reduce")
    dp.close()
    # --- single process on the concatenated batch (every rank computes it; rank 0 reports)
    s_1, t_1 = models()
    tr1 = SSLTrainer(s_1, t_1, n_classes=k, threshold=0.6)
    tr1.iter_num = 3000
    outs1 = [tr1.step({**{kk: v.cuda() for kk, v in full.items()}, **extra}) for _ in range(2)]
    torch.cuda.synchronize()
    def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    worst = max(rel(p, q) for p, q in zip(s_dp.parameters(), s_1.parameters()))
    worst_t = max(rel(p, q) for p, q in zip(t_dp.parameters(), t_1.parameters()))
    worst_rs = max(rel(a, b) for (n1, a), (n2, b) in zip(s_dp.named_buffers(), s_1.named_buffers()) if "running" in n1)
    dl = [abs(float(o["loss"]) - float(o1["loss"])) / abs(float(o1["loss"])) for o, o1 in zip(outs, outs1)]
    same = all(torch.equal(outs[0][kk], outs1[0][kk][sl]) for kk in ("pseudo_label", "mask_w"))
    gathered = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([worst], device="cuda"))
    if rank == 0:
        print(f"dp_check world={world} precision={precision} model={'unet_b dsbn x3' if dsbn else 'unet_a'}: loss rel diff per step {dl}, student weights {worst:.2e} (all ranks {[float(g) for g in gathered]}), "
              f"teacher {worst_t:.2e}, running stats {worst_rs:.2e}, step-0 masks identical {same}")
        cat = lambda m: torch.cat([q.detach().double().flatten() for q in m.parameters()])
        allw = rel(cat(s_dp), cat(s_1))
        print(f"all student weights concatenated: {allw:.2e}")
        # per-tensor: BN biases start at 0, so after two steps they ARE the (ill-conditioned, ReLU-flip-prone) gradient: 5e-2
        tol = 2e-4 if precision == "fp32" else 5e-2
        wtol = 5e-2 if precision == "fp32" else 0.3          # per tensor (BatchNorm biases start at 0: after two steps they ARE the gradient)
        if world > 2:     # 64 x 64 inputs: the bottleneck BatchNorms see 16 x world pixels; summation order matters more as ranks are added
            tol, wtol = max(tol, 2e-3), 1.0
        if dsbn:          # UNet-B gradients are ill-conditioned even between two fp32 evaluations (tests/test_parity_fullsize_gpu.py)
            tol, wtol = (1e-3, 0.25) if precision == "fp32" else (5e-2, 0.6)
        assert allw < tol and worst < wtol and worst_t < wtol and worst_rs < tol and max(dl) < tol, "data-parallel step != single-process step on the concatenated batch"
        print("dp_check OK")
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
