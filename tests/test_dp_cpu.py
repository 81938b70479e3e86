"""world_size-2 gloo tests (CPU) of the data-parallel host logic in ustrun/dp.py: gradient buckets
over the flat buffer (launch order, overlap bookkeeping, averaging) and the cross-rank BatchNorm
statistics hook.  The compute on each rank is the CPU oracle; the collectives are real."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _buckets_case(rank, world):
    from ustrun.dp import GradBuckets
    sizes = [7, 64, 3, 129, 1000, 5, 256]
    offsets, total = [], 0
    for n in sizes:
        offsets.append(total)
        total += (n + 3) // 4 * 4
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    local = flat.clone()
    b = GradBuckets(flat, offsets, sizes, bucket_bytes=512)
    assert len(b.ranges) >= 3 and b.ranges[0][0] == 0 and b.ranges[-1][1] == total
    assert all(b.ranges[i][1] == b.ranges[i + 1][0] for i in range(len(b.ranges) - 1)), "buckets must tile the buffer"
    launched_before_flush = 0
    for i in reversed(range(len(sizes))):          # backward touches tensors in reverse order
        b.mark(i)
        launched_before_flush = sum(b.launched)
    assert launched_before_flush == len(b.ranges), "every bucket goes as soon as its tensors are marked"
    b.flush()
    # second step: nothing marked (e.g. world of skipped layers) -> flush launches everything
    flat2 = flat.clone()
    b2 = GradBuckets(flat2, offsets, sizes, bucket_bytes=1 << 30)
    assert len(b2.ranges) == 1
    b2.flush()
    return local, flat, flat2


def test_grad_buckets_allreduce_sum():
    out = _spawn(_buckets_case)
    want = out[0][0] + out[1][0]
    for r in (0, 1):
        assert torch.allclose(out[r][1], want) and torch.allclose(out[r][2], 2 * want - want + want)  # flat2 = allreduce(allreduced)


def _syncbn_case(rank, world):
    """Global-batch BN statistics from per-rank partial sums == statistics of the concatenated batch."""
    from ustrun import bridge
    from ustrun.dp import DataParallel
    dp = DataParallel(sync_bn=True)
    assert bridge.BN_SYNC is not None and bridge.BN_WORLD == world
    g = torch.Generator().manual_seed(7)
    full = torch.randn(4, 8, 6, 6, generator=g)
    mine = full[rank * 2:(rank + 1) * 2]
    sums = torch.cat([mine.sum((0, 2, 3)), (mine * mine).sum((0, 2, 3))])
    bridge.BN_SYNC(sums)
    n = full.numel() / 8
    mean, var = sums[:8] / n, sums[8:] / n - (sums[:8] / n) ** 2
    dp.close()
    assert bridge.BN_SYNC is None
    return mean, var, full.mean((0, 2, 3)), full.var((0, 2, 3), unbiased=False)


def test_sync_bn_statistics():
    out = _spawn(_syncbn_case)
    for r in (0, 1):
        mean, var, ref_mean, ref_var = out[r]
        assert torch.allclose(mean, ref_mean, atol=1e-6) and torch.allclose(var, ref_var, atol=1e-5)


def _dp_step_case(rank, world):
    """Each rank: oracle loss/grad on its half of the batch -> DataParallel hooks (begin_step,
    on_grad_done in backward order during the last branch -- each AFTER its gradient was written, with a bucket size that
    puts single tensors into their own buckets, the case the round-1 advisor flagged -- finish_step) -> averaged gradient."""
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun.dp import DataParallel

    class FakeOpt:       # the slice of FusedSGDEMA that DataParallel touches, on CPU tensors
        def __init__(self, params):
            self.params = params
            self.offsets, tot = [], 0
            for p in params:
                self.offsets.append(tot)
                tot += (p.numel() + 3) // 4 * 4
            self.flat_grad = torch.zeros(tot)

        def view(self, i):
            return self.flat_grad[self.offsets[i]: self.offsets[i] + self.params[i].numel()].view(self.params[i].shape)

    st = U.init_unet_b(1, 2, n=4, seed=3)
    params, _ = U.split_state(st)
    plist = list(params.values())
    for p in plist:
        p.requires_grad_(True)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 1, 32, 32, generator=g)
    t = torch.randint(0, 2, (4, 32, 32), generator=g)
    sl = slice(rank * 2, rank * 2 + 2)
    loss = S.masked_term(U.unet_b_forward(st, x[sl], True), t[sl], None, 2, "softmax")
    loss.backward()
    opt = FakeOpt(plist)
    assert DataParallel(sync_bn=False, global_loss=True).finish_step.__name__ == 'finish_step'
    dp = DataParallel(sync_bn=False, global_loss=False, bucket_bytes=64)       # per-rank losses -> averaged gradients
    dp.begin_step(opt)
    assert len(dp.buckets.ranges) > len(plist) // 2
    for i in reversed(range(len(plist))):
        opt.view(i).copy_(plist[i].grad)
        dp.on_grad_done(plist[i], last_branch=True)          # the bucket may go NOW: its gradient must already be in place
    assert all(dp.buckets.launched), "every bucket was launched by its last member"
    scale = dp.finish_step(opt)
    assert scale == 0.5
    return opt.flat_grad * scale, [p.grad.clone() for p in plist], opt.offsets


def test_dp_gradient_average_matches_single_process():
    out = _spawn(_dp_step_case)
    flat0, grads0, offs = out[0]
    flat1, grads1, _ = out[1]
    assert torch.equal(flat0, flat1), "all ranks end with the same averaged gradient"
    for i, (a, b) in enumerate(zip(grads0, grads1)):
        want = (a + b) / 2
        got = flat0[offs[i]: offs[i] + a.numel()].view(a.shape)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-8)


def test_peer_barrier_sequence_numbers():
    """csrc/peer_bn.cu: seq_eff = (seq - 1 + *seq_base) % 0x7FFFFFFE + 1 in uint32 arithmetic, seq = index within the step
    (launch argument, fixed under CUDA-graph replay), *seq_base = barriers of earlier steps modulo 0x7FFFFFFE (ustrun.dp.PeerStats.begin_step).
    The protocol needs: never 0, consecutive barriers alternate parity (two packet slots), also across the wrap-around."""
    from ustrun.dp import SEQ_MOD, barrier_seq

    def device(seq, base):           # the kernel's 32-bit arithmetic
        assert 0 <= base < SEQ_MOD and 1 <= seq < (1 << 20)
        x = (seq - 1 + base) & 0xFFFFFFFF
        assert x == seq - 1 + base   # no 32-bit overflow before the modulo
        return x % SEQ_MOD + 1

    per_step = 234
    for total in (0, 1, per_step, 5 * per_step, SEQ_MOD - 300, SEQ_MOD - 1, SEQ_MOD, 3 * SEQ_MOD + 17):
        prev = None
        for step in range(3):
            base = (total + step * per_step) % SEQ_MOD
            for idx in range(1, per_step + 1):
                s = device(idx, base)
                assert s == barrier_seq(total + step * per_step, idx)
                assert 1 <= s <= SEQ_MOD
                if prev is not None:
                    assert (s & 1) != (prev & 1) and s != prev
                prev = s
