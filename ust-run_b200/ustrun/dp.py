"""Data parallelism for the SSL step: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the plumbing.  The reference is single-GPU, so this layer is new:

  * gradients: the fused optimiser keeps every gradient in ONE flat fp32 buffer; ``GradBuckets``
    cuts it into contiguous buckets and all-reduces each bucket on a side stream as soon as the
    LAST loss branch's backward has written all of its tensors (reverse layer order), overlapping
    the remaining backward.  The 1/world averaging is folded into the SGD kernel (``grad_scale``).
  * loss: the CE/Dice partial sums (3C+1 floats per term) are summed across ranks before the loss
    finalize kernel (``global_loss``), so G ranks x B images reproduce one device with G*B images.
  * BatchNorm / DSBN statistics: per layer, per pass, a [2*C] fp32 vector (sum, sum of squares;
    backward: sum g', sum g' xhat) is summed across ranks before the finalize kernel, giving
    global-batch BN (G ranks x B images == one device with G*B images).
The host-side logic is device-agnostic so that it is covered by world_size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

from typing import List, Optional

import os

import torch
import torch.distributed as dist

from . import bridge


class GradBuckets:
    """Contiguous buckets over a flat gradient buffer; ``mark(i)`` says tensor i's gradient kernel has
    been enqueued, a bucket is launched when all of its tensors are marked."""

    def __init__(self, flat: torch.Tensor, offsets: List[int], sizes: List[int], bucket_bytes: int = 32 << 20, group=None):
        self.flat, self.group = flat, group
        self.offsets, self.sizes = offsets, sizes
        self.bucket_of, self.ranges = [], []
        start, cur = 0, 0
        n = len(offsets)
        for i in range(n):
            end = offsets[i + 1] if i + 1 < n else flat.numel()
            self.bucket_of.append(len(self.ranges))
            cur = end - start
            if cur * 4 >= bucket_bytes or i == n - 1:
                self.ranges.append((start, end))
                start = end
        self.members = [[i for i in range(n) if self.bucket_of[i] == b] for b in range(len(self.ranges))]
        self.comm_stream = torch.cuda.Stream() if flat.is_cuda else None
        self.reset()

    def reset(self):
        self.pending = [set(m) for m in self.members]
        self.launched = [False] * len(self.ranges)
        self.streams = [dict() for _ in self.ranges]
        self.works = []

    def mark(self, i: int, stream=None):
        """``stream``: the stream tensor i's gradient kernel was enqueued on (weight gradients run on the
        engine's side stream, BatchNorm parameter gradients on the main one); the bucket waits for every
        stream that produced one of its members."""
        b = self.bucket_of[i]
        self.pending[b].discard(i)
        if stream is not None:
            self.streams[b][stream.cuda_stream] = stream
        if not self.pending[b] and not self.launched[b]:
            self._launch(b)

    def _launch(self, b: int):
        s, e = self.ranges[b]
        self.launched[b] = True
        view = self.flat[s:e]
        if self.comm_stream is not None:
            producers = list(self.streams[b].values()) or [torch.cuda.current_stream()]
            events = []
            for st in producers:
                ev = torch.cuda.Event()
                ev.record(st)
                events.append(ev)
            with torch.cuda.stream(self.comm_stream):
                for ev in events:
                    self.comm_stream.wait_event(ev)
                self.works.append(dist.all_reduce(view, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(view, group=self.group, async_op=True))

    def flush(self):
        """Launch whatever is still pending and make the current stream wait for all buckets."""
        for b in range(len(self.ranges)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.works:
            w.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.works = []


SEQ_MOD = 0x7FFFFFFE          # even: consecutive barriers alternate slot parity across the wrap-around as well


def barrier_seq(total: int, index: int) -> int:
    """Sequence number of the ``index``-th (1-based) cross-rank BatchNorm barrier of a step that follows ``total`` completed
    barriers: host-side statement of what csrc/peer_bn.cu computes from its launch argument ``seq = index`` and the device word
    ``*seq_base = total % SEQ_MOD`` in 32-bit arithmetic.  1 .. 2^31-2, never 0 (the packet buffers start zeroed)."""
    return (total + index - 1) % SEQ_MOD + 1


class PeerStats:
    """Symmetric (peer-mapped) buffers for the BatchNorm finalize kernels that reduce their [2*C]
    statistics across ranks themselves over NVLink (csrc/peer_bn.cu) instead of calling NCCL."""

    def __init__(self, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm

        from . import _lib as L
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = int(L.lib.ustrun_peer_buffer_bytes())
        self.buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, grp)
        self._bases = (ctypes.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        # sequence numbers: barriers of completed steps (host `total`, device word `base` = total mod 0x7FFFFFFE, advanced by
        # begin_step) + the call's index within the step (launch argument): see include/ustrun.h
        self.base = torch.zeros(1, dtype=torch.int32, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        self.total = 0
        self.calls = 0
        self._ring = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(4)]
        self._ring_ev = [None] * 4
        self._ring_i = 0
        self._ctypes = ctypes
        self._poll_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._poll_ev = None
        torch.cuda.synchronize()
        dist.barrier(grp)

    SEQ_MOD = SEQ_MOD

    @property
    def seq(self):
        """Barriers issued so far."""
        return self.total + self.calls

    def begin_step(self):
        """Once per step, on the stream the step is about to run on and OUTSIDE any graph capture: fold the previous step's
        calls into the device-side base.  The launch arguments of the step's finalize kernels (1, 2, 3, ...) then repeat from
        step to step -- what a CUDA-graph replay needs."""
        self.total += self.calls
        self.calls = 0
        i = self._ring_i & 3
        self._ring_i += 1
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()            # the copy that last read this pinned word (four steps ago) is done
        self._ring[i][0] = self.total % self.SEQ_MOD
        self.base.copy_(self._ring[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev

    def args(self):
        """(peer_bases, rank, world, seq, seq_base, error) for ustrun_bn_*_finalize_peer; every rank
        issues the same sequence of calls, so the sequence numbers agree."""
        self.calls += 1
        c = self._ctypes
        return (self._bases, self.rank, self.world, self.calls, c.c_void_p(self.base.data_ptr()), c.c_void_p(self.error.data_ptr()))

    def check(self):
        if int(self.error.item()) != 0:
            raise RuntimeError("peer BatchNorm reduction timed out waiting for another rank")

    def poll(self):
        """Non-blocking check of the device-side timeout flag: raises if a finalize kernel of an EARLIER step gave up waiting
        for a peer (it then continued with this rank's local sums, i.e. unsynchronised statistics).  An asynchronous copy of
        the flag is enqueued on every call and read on the next one, so a timeout surfaces at most two steps late without
        ever synchronising the stream."""
        if self._poll_ev is not None and self._poll_ev.query():
            if int(self._poll_host[0]) != 0:
                raise RuntimeError("peer BatchNorm reduction timed out waiting for another rank: BatchNorm statistics of that step "
                                   "were not synchronised (a rank stalled for longer than the kernel's timeout)")
            self._poll_ev = None
        if self._poll_ev is None:
            self._poll_host.copy_(self.error, non_blocking=True)
            self._poll_ev = torch.cuda.Event()
            self._poll_ev.record()


class _BNSync:
    """bridge.BN_SYNC object: callable NCCL fallback + optional fused peer path."""

    def __init__(self, group, peer):
        self.group, self.peer = group, peer

    def __call__(self, t: torch.Tensor):
        dist.all_reduce(t, group=self.group)


class DataParallel:
    """Attach to an ``SSLTrainer`` (``dp=`` argument).  ``sync_bn`` installs the cross-rank
    BatchNorm statistics hook used by every conv+BN op."""

    def __init__(self, group=None, bucket_bytes: int = 32 << 20, sync_bn: bool = True, global_loss: bool = True, force: bool = False):
        assert dist.is_initialized(), "init_process_group first (backend nccl on GPUs, gloo in CPU tests)"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # force: run every collective / peer kernel even in a 1-rank group (single-GPU tests of this code path)
        self.active = self.world > 1 or bool(force)
        self.bucket_bytes = bucket_bytes
        self.buckets: Optional[GradBuckets] = None
        self.branch = 0
        self.sync_bn = sync_bn
        # global_loss: the (3C+1) CE/Dice partial sums are summed across ranks inside every loss term, so
        # the step equals the single-device step on the concatenated batch (gradients are then SUMMED);
        # otherwise each rank has its own loss and gradients are averaged (plain DDP semantics).
        self.global_loss = global_loss
        self.peer = None
        self.graph_safe = False          # set below: replayable as a CUDA graph with the peer-memory BN path or without sync BN
        if sync_bn and self.active:
            # sync_bn="peer" (default on GPUs): finalize kernels reduce over NVLink peer memory themselves;
            # sync_bn="nccl": one NCCL all-reduce per layer and pass (validation / CPU tests)
            want_peer = sync_bn == "peer" or (sync_bn is True and torch.cuda.is_available() and dist.get_backend(group) == "nccl")
            if want_peer:
                try:
                    self.peer = PeerStats(group)
                except Exception as e:          # symmetric memory unavailable on this platform: say so, use NCCL
                    if sync_bn == "peer":
                        raise
                    print(f"[ustrun.dp] peer-memory BatchNorm statistics unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
            bridge.BN_SYNC = _BNSync(group, self.peer)
            bridge.BN_WORLD = self.world
        # CUDA-graph replay of the whole data-parallel step: the bucket all-reduces and the loss-sum all-reduces are captured
        # NCCL calls; the peer-memory BN kernels take their sequence number from a device word.  The per-layer NCCL BN path
        # (validation) stays eager.
        self.graph_safe = torch.cuda.is_available() and (self.peer is not None or not (sync_bn and self.active)) \
            and os.environ.get("USTRUN_DP_GRAPH", "1") != "0"

    def close(self):
        bridge.BN_SYNC, bridge.BN_WORLD = None, 1

    def sum_across_ranks(self, t: torch.Tensor):
        dist.all_reduce(t, group=self.group)

    def broadcast_parameters(self, tensors):
        """Rank 0's tensors to every rank (called by SSLTrainer at attach time: weights and BatchNorm buffers of both models)."""
        for t in tensors:
            dist.broadcast(t.data if hasattr(t, "data") else t, src=0, group=self.group)

    # -- hooks called by SSLTrainer ----------------------------------------------------------
    def begin_step(self, opt):
        if self.buckets is None or self.buckets.flat is not opt.flat_grad:
            sizes = [p.numel() for p in opt.params]
            self.buckets = GradBuckets(opt.flat_grad, opt.offsets, sizes, self.bucket_bytes, self.group)
            self._index = {id(p): i for i, p in enumerate(opt.params)}
        self.buckets.reset()
        self.branch = 0

    def on_branch_done(self):
        self.branch += 1

    def on_grad_done(self, param, last_branch: bool):
        """Called right AFTER the kernel that writes ``param``'s gradient (the last loss branch's contribution) has been
        enqueued on the current stream: its bucket may be all-reduced once every member has reported, ordered after the
        streams that produced them (weight gradients: the engine's side stream; BatchNorm / bias gradients: the main one)."""
        if not last_branch or not self.active:
            return
        cur = torch.cuda.current_stream() if torch.cuda.is_available() else None
        self.buckets.mark(self._index[id(param)], cur)

    def finish_step(self, opt) -> float:
        if self.active:
            self.buckets.flush()
        return 1.0 if self.global_loss else 1.0 / self.world
