"""Data-parallel code path on ONE GPU (the driver's GPU tier has a single device): a world-size-1 NCCL process group drives
``ustrun.dp.DataParallel`` end to end -- gradient buckets on the communication stream (tiny buckets: one tensor each),
global-loss partial sums, and the PEER-MEMORY BatchNorm finalize kernels of csrc/peer_bn.cu (symmetric-memory buffer, flags,
sequence numbers; with one rank the cross-rank sum has a single term) -- and must reproduce the plain single-GPU step.
Multi-rank parity (2 ranks == one process on the concatenated batch, incl. DSBN) is tools/dp_check.py under
``gpurun --gpus 2`` and bench.py's ``dp_parity`` record; the host logic runs under gloo in tests/test_dp_cpu.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def nccl_world1():
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), RANK="0", WORLD_SIZE="1")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    import gc
    gc.collect()                     # step graphs that captured NCCL kernels must be gone before the communicator
    torch.cuda.synchronize()
    dist.destroy_process_group()


def _pair(kind, **kw):
    torch.manual_seed(1337)
    if kind == "a":
        from networks.unet_model import UNet
        s, t = UNet(1, 2, **kw), UNet(1, 2, **kw)
    else:
        from networks.unet import UNet
        s, t = UNet(3, 2, **kw), UNet(3, 2, **kw)
    t.load_state_dict(s.state_dict())
    for p in t.parameters():
        p.detach_()
    return s.cuda().train(), t.cuda().train()


@pytest.mark.parametrize("kind,sync", [("a", "peer"), ("a", "nccl"), ("b_dsbn", "peer")])
def test_world1_data_parallel_equals_plain_step(nccl_world1, kind, sync):
    from ustrun import bridge
    from ustrun import engine as E
    from ustrun import synth as S
    from ustrun.dp import DataParallel
    from ustrun.step import SSLTrainer
    E.set_precision("fp32")          # the two paths sum the same statistics in different orders: compare at fp32 resolution
    dsbn = kind == "b_dsbn"
    kw = dict(norm="dsbn", num_domains=3) if dsbn else {}
    c = 3 if dsbn else 1
    extra = dict(domain_lb=1, domain_ulb=2) if dsbn else {}
    batches = [{kk: v.cuda() for kk, v in S.synthetic_batch(c, 2, 64, 64, 2, 2, seed=30 + i).items()} for i in range(3)]
    s0, t0 = _pair(kind[0], **kw)
    plain = SSLTrainer(s0, t0, n_classes=2, threshold=0.6)
    ref = [plain.step({**b, **extra})["loss"].clone() for b in batches]
    E.set_precision("fp32")
    s1, t1 = _pair(kind[0], **kw)
    try:
        dp = DataParallel(sync_bn=sync, global_loss=True, bucket_bytes=256, force=True)
    except Exception as e:                                   # symmetric memory not available for a 1-rank group on this box
        if sync == "peer":
            pytest.skip(f"peer-memory buffers unavailable: {type(e).__name__}: {e}")
        raise
    try:
        assert (dp.peer is not None) == (sync == "peer")
        tr = SSLTrainer(s1, t1, n_classes=2, threshold=0.6, dp=dp)
        got = [tr.step({**b, **extra})["loss"].clone() for b in batches]
        torch.cuda.synchronize()
        if dp.peer is not None:
            dp.peer.check()
            assert dp.peer.seq > 100                          # every BatchNorm layer of every pass went through the peer kernels
    finally:
        dp.close()
        E.set_precision("bf16")
        assert bridge.BN_SYNC is None
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 1e-3 * abs(float(b)), (float(a), float(b))
    cat = lambda m: torch.cat([p.detach().double().flatten() for p in m.parameters()])
    # same maths, different summation trees (partial rows -> float sums -> finalize, bucketed gradients): fp32 noise through
    # three steps of an ill-conditioned random-init network
    assert float((cat(s1) - cat(s0)).norm() / cat(s0).norm()) < 5e-3
    for (n, a), (_, b) in zip(s1.named_buffers(), s0.named_buffers()):
        assert float((a.double() - b.double()).norm() / (b.double().norm() + 1e-12)) < 5e-3, n


@pytest.mark.parametrize("kind", ["a", "b_dsbn"])
def test_world1_data_parallel_lanes_bit_identical(nccl_world1, kind):
    """Multi-lane step under data parallelism with the peer-memory BatchNorm path: every cross-rank finalize kernel runs on the
    ONE BN-sync stream in program order (engine.on_sync_stream), running statistics are deferred through the kernel's
    ``stat_out`` -- the result must equal the single-lane data-parallel step bit for bit.  So must the CUDA-graph replay of the
    data-parallel step (sequence numbers of the barriers from the device-side base, include/ustrun.h)."""
    from ustrun import bridge
    from ustrun import engine as E
    from ustrun import synth as S
    from ustrun.dp import DataParallel
    from ustrun.step import SSLTrainer
    dsbn = kind == "b_dsbn"
    kw = dict(norm="dsbn", num_domains=3) if dsbn else {}
    c = 3 if dsbn else 1
    extra = dict(domain_lb=1, domain_ulb=2) if dsbn else {}
    batches = [{kk: v.cuda() for kk, v in S.synthetic_batch(c, 2, 64, 64, 2, 2, seed=60 + i).items()} for i in range(6)]
    runs = []
    for lanes, graph in ((1, False), (3, False), (2, True)):
        E.set_precision("bf16")
        s, t = _pair(kind[0], **kw)
        try:
            dp = DataParallel(sync_bn="peer", global_loss=True, bucket_bytes=256, force=True)
        except Exception as e:
            pytest.skip(f"peer-memory buffers unavailable: {type(e).__name__}: {e}")
        try:
            assert dp.graph_safe
            tr = SSLTrainer(s, t, n_classes=2, threshold=0.6, dp=dp, lanes=lanes, use_graph=graph)
            assert tr.lanes == lanes
            losses = [tr.step({**b, **extra}, lq=b["ulb_w"][:1].contiguous())["loss"].clone() for b in batches]
            torch.cuda.synchronize()
            dp.peer.check()
            if graph:       # steps 3.. were replays of ONE captured graph (NCCL bucket / loss-sum all-reduces and peer-BN kernels inside)
                assert tr.use_graph and len(tr._graphs) == 1, getattr(tr, "graph_error", None)
            assert dp.peer.seq == runs[0][3] if runs else True        # same number of cross-rank barriers whichever way the step ran
            nseq = dp.peer.seq
            tr._graphs.clear()       # captured NCCL kernels: released before the process group (see bench.py::shutdown_dp)
        finally:
            dp.close()
            assert bridge.BN_SYNC is None
        runs.append((losses, s, t, nseq))
    l1, s1, t1, _ = runs[0]
    for lx, sx, tx, _ in runs[1:]:
        for a, b in zip(l1, lx):
            assert float(a) == float(b)
        for m1, mx in ((s1, sx), (t1, tx)):
            for (n, a), (_, b) in zip(list(m1.named_parameters()) + list(m1.named_buffers()), list(mx.named_parameters()) + list(mx.named_buffers())):
                assert torch.equal(a, b), n
