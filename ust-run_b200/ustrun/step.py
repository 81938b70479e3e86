"""The fused SSL training step (Tier B of SURVEY 8b): train.py:638-856 / train_mnms.py:586-771 as
one call that launches only libustrun_sm100.so kernels and never synchronises with the host.

    trainer = SSLTrainer(model, ema_model, n_classes=2, branch="softmax", base_lr=0.03, max_iterations=60000)
    out = trainer.step(batch)        # batch: dict of CUDA tensors, see ``SSLTrainer.step``

Order of work (identical to the reference, including the order in which the student's BatchNorm
running statistics are updated: ulb_w, lb_x, s_ul, s_lu, ulb_s, [lq_s]):
  1. teacher (train-mode BN, no grad) on ulb_w, CutMix(ulb_w, mix) and CutMix(mix, ulb_w)
  2. student forward on ulb_w (no grad; only feeds the hardness statistic)           train.py:668
  3. ONE fused pseudo-label kernel -> 9 uint8 planes                                train.py:649-697
  4. for each of the four loss branches: forward -> fused CE+Dice pass 1 -> pass 2 (dLoss/dlogits
     scaled by 1, cw, cw, cw^2) -> backward, accumulating weight gradients in one flat fp32
     buffer.  The loss is a sum of terms that each depend on one forward and on constant pseudo
     labels, so branch-at-a-time is mathematically the reference's single backward
     while holding one branch of activations instead of five.
  5. [data parallel: bucketed NCCL all-reduce of the flat gradient buffer]
  6. ONE fused SGD(momentum, wd) + EMA kernel over all tensors                      train.py:848-851
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from . import bridge
from . import engine as E
from .engine import _call, _profiled, _ptr, _stream
from .loss_ops import as_u8, term_backward, term_forward
from .optim import FusedSGDEMA


def sigmoid_rampup(current, rampup_length):
    """utils/ramps.py:19-26 (numpy double arithmetic, bit-identical to the reference)."""
    if rampup_length == 0:
        return 1.0
    phase = 1.0 - float(np.clip(current, 0.0, rampup_length)) / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def pseudo_labels(t1, t2, t3, box, cut_label, cut_mask, choice, threshold, branch, student_logits=None) -> Dict[str, torch.Tensor]:
    """Fused pseudo-label / ensemble / CutMix-label composition (train.py:649-697).

    t1,t2,t3: teacher logits fp32 NCHW for ulb_w, CutMix(ulb_w, mix), CutMix(mix, ulb_w);
    box [Bu,H,W] in {0,1}; cut_label/cut_mask: labelled batch (+ bank) labels/masks; choice [Bu].
    Returns uint8 planes ([Bu,H,W] softmax branch, [Bu,C,H,W] sigmoid branch)."""
    L.require_device()
    Bu, C, H, W = t1.shape
    dev = t1.device
    t1, t2, t3 = (t.float().contiguous() for t in (t1, t2, t3))
    s0 = None if student_logits is None else student_logits.float().contiguous()
    box_u8 = as_u8(box).reshape(Bu, H, W).contiguous()
    choice_i = choice.to(device=dev, dtype=torch.int32).contiguous()
    shape = (Bu, H, W) if branch == "softmax" else (Bu, C, H, W)
    cl = as_u8(cut_label).reshape((-1,) + shape[1:]).contiguous()
    cm = as_u8(cut_mask).reshape((-1,) + shape[1:]).contiguous()
    names = ["pseudo_label", "mask", "pseudo_label_w", "mask_w", "pseudo_label_ul", "mask_ul", "pseudo_label_lu", "mask_lu", "stu_pseudo_label"]
    planes = torch.empty((9,) + shape, dtype=torch.uint8, device=dev)
    outs = [planes[i] for i in range(9)]
    ptrs = [_ptr(o) for o in outs]
    if s0 is None:
        ptrs[8] = None
    # algorithmic bytes: 4 logit tensors (fp32) + box / cut label / cut mask read, 9 uint8 planes written
    nlab = Bu * H * W * (1 if branch == "softmax" else C)
    pl_bytes = (4 if s0 is not None else 3) * Bu * C * H * W * 4 + Bu * H * W + 2 * nlab + (9 if s0 is not None else 8) * nlab
    if branch == "softmax":
        _profiled("hbm_pseudo_label", pl_bytes, "ustrun_pseudo_label_softmax", _ptr(t1), _ptr(t2), _ptr(t3), _ptr(s0), _ptr(box_u8), _ptr(cl), _ptr(cm), _ptr(choice_i),
              float(threshold), Bu, C, H, W, *ptrs, _stream())
    else:
        _profiled("hbm_pseudo_label", pl_bytes, "ustrun_pseudo_label_sigmoid", _ptr(t1), _ptr(t2), _ptr(t3), _ptr(s0), _ptr(box_u8), _ptr(cl), _ptr(cm), _ptr(choice_i),
              float(threshold), float(1 - threshold), Bu, C, H, W, *ptrs, _stream())
    res = dict(zip(names, outs))
    if s0 is None:
        del res["stu_pseudo_label"]
    return res


def normalize_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """Normalize_tf + ToTensor on the device (custom_transforms.py:650-684, 728-753): uint8 [B,H,W,C] (or [B,H,W]) -> float32
    [B,C,H,W] in [-1, 1], bit-exact with the reference's numpy arithmetic."""
    L.require_device()
    x = images_u8.contiguous()
    if x.dim() == 3:
        x = x.unsqueeze(-1)
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    _call("ustrun_normalize_u8_to_nchw", _ptr(x), _ptr(out), B, C, H, W, _stream())
    return out


def _is_u8_image(t):
    return t is not None and t.dtype == torch.uint8


def mix_input(a, b, box_u8, b_index=None) -> E.Act:
    """NHWC activation of ``a*(1-box) + b[b_index]*box``; box=None converts a.  Sources are float32 NCHW tensors in [-1, 1]
    or uint8 [B,H,W,C] image batches as the loaders hold them before Normalize_tf + ToTensor (normalised on the fly)."""
    if not (_is_u8_image(a) or _is_u8_image(b)):
        a = a.float().contiguous()
        B, C, H, W = a.shape
        out = E.Act.new(B, H, W, C, device=a.device)
        _call("ustrun_mix_to_nhwc", _ptr(a), _ptr(b.float().contiguous()) if b is not None else None,
              _ptr(b_index) if b_index is not None else None, _ptr(box_u8) if box_u8 is not None else None, out.ptr, out.ld, out.dtype_code,
              B, C, H, W, _stream())
        return out
    au = a.contiguous() if _is_u8_image(a) else None
    af = None if au is not None else a.float().contiguous()
    bu = b.contiguous() if _is_u8_image(b) else None
    bf = None if (b is None or bu is not None) else b.float().contiguous()
    if au is not None:
        B, H, W, C = au.shape if au.dim() == 4 else (au.shape + (1,))
    else:
        B, C, H, W = af.shape
    out = E.Act.new(B, H, W, C, device=a.device)
    _call("ustrun_mix_any_to_nhwc", _ptr(af), _ptr(au), _ptr(bf), _ptr(bu), _ptr(b_index) if b_index is not None else None,
          _ptr(box_u8) if box_u8 is not None else None, out.ptr, out.ld, out.dtype_code, B, C, H, W, _stream())
    return out


STEP_DOMAIN_TAGS = ("t1", "t2", "t3", "s0", "lb", "ul", "lu", "s", "lq")
# hyper block layout (float32, device): the per-step scalars every kernel of the step reads from memory
H_LR, H_ALPHA, H_GSCALE, H_CW, H_CW2, H_ONE = 0, 1, 2, 3, 4, 5
_HYPER_N = 8
_HYPER_RING = 4


def step_domains(domain_lb, domain_ulb):
    """Domain label of each forward of a DSBN step.  Upstream has no DSBN step; the convention here: a forward
    belongs to the domain of the image that fills the area OUTSIDE the CutMix box -- t1 / t2 / s0 / ul / s / lq: the
    unlabelled batch's domain, t3 / lb / lu: the labelled batch's.  DSBN semantics are the reference's (dsbn.py:24-27):
    the whole forward batch goes through ``bns[domain_label[0]]``."""
    u, l = [int(domain_ulb)], [int(domain_lb)]
    return dict(t1=u, t2=u, t3=l, s0=u, lb=l, ul=u, lu=l, s=u, lq=u)


class SSLTrainer:
    def __init__(self, model, ema_model, n_classes, branch="softmax", base_lr=0.03, max_iterations=30000, threshold=0.95,
                 consistency=1.0, consistency_rampup=200.0, ema_decay=0.99, momentum=0.9, weight_decay=1e-4, dp=None,
                 forward_kwargs=None, fft_window=0.01, hardness_mode=None, use_graph=False, lanes=1):
        self.model, self.ema_model = model, ema_model
        self.n_classes, self.branch = n_classes, branch
        self.base_lr, self.lr, self.max_iterations = base_lr, base_lr, max_iterations
        self.threshold, self.consistency, self.consistency_rampup, self.ema_decay = threshold, consistency, consistency_rampup, ema_decay
        self.iter_num = 0
        self.fft_window = fft_window          # --LB (train.py:76): half-width of the amplitude window as a fraction of min(H, W)
        # "binary" | "2label" | "3label": also return the per-sample hardness / hardest-sample index (train.py:705-718)
        self.hardness_mode = hardness_mode
        self.first_epoch = False              # train.py:711-713: hardness is 1 for every sample during epoch 0
        self.dp = dp
        self.forward_kwargs = forward_kwargs or {}
        self.params = list(model.parameters())
        if dp is not None and getattr(dp, "active", False) and getattr(dp, "broadcast_at_attach", True):
            # every rank must start from rank 0's weights AND BatchNorm buffers (the reference seeds one process; here each
            # rank constructs its own modules): one broadcast per tensor, once
            with torch.no_grad():
                dp.broadcast_parameters([t for m in (model, ema_model) for t in list(m.parameters()) + list(m.buffers())])
        self.opt = FusedSGDEMA(self.params, list(ema_model.parameters()), momentum=momentum, weight_decay=weight_decay)
        self._touched = set()
        self._last_branch = False
        dev = self.params[0].device
        # per-step scalars live in device memory (lr, EMA alpha, gradient scale, consistency weights): the launch arguments of
        # a step never change, which is what lets the whole step be captured once as a CUDA graph and replayed
        self.hyper = torch.zeros(_HYPER_N, dtype=torch.float32, device=dev)
        self._hyper_host = [torch.zeros(_HYPER_N, dtype=torch.float32).pin_memory() for _ in range(_HYPER_RING)]
        self._hyper_ev = [None] * _HYPER_RING
        self._hyper_slot = 0
        self.use_graph = bool(use_graph)
        # lanes > 1: the independent forwards of a step (t1 | t2 | t3 | s0, then the loss branches lb | ul | lu | s | lq) are
        # enqueued round-robin on `lanes` CUDA streams.  Results are bit-identical to lanes = 1: BatchNorm running statistics
        # are logged per forward and applied in the reference's order at the end of the step (E.StatLog), and every
        # parameter-gradient accumulation runs on the one weight-gradient stream in program order.
        # Data parallel: lanes are allowed with the peer-memory BatchNorm path (every cross-rank kernel then runs on ONE
        # stream in program order, engine.on_sync_stream) and without cross-rank statistics; not with per-layer NCCL calls.
        dp_ok = dp is None or not getattr(dp, "active", False) or not dp.sync_bn or getattr(dp, "peer", None) is not None
        self.lanes = max(1, int(lanes)) if dp_ok else 1
        self._lane_streams = []
        self._stats = E.StatLog() if self.lanes > 1 else None
        self._pack_plan = E.PackPlan(E.collect_packed(model, ema_model))
        self._graphs = {}
        self._eager_steps = 0
        self._eager_by_key = {}
        self._copy_stream, self._up_slot, self._staging, self._consumed = None, 0, [{}, {}], [None, None]
        self.launches_per_step = 0            # kernels of the last eager / captured step (bench.py: gpu_launches)
        self._domains = None

    # -- helpers ---------------------------------------------------------------------------
    def consistency_weight(self, iter_num):
        return self.consistency * sigmoid_rampup(iter_num // (self.max_iterations / self.consistency_rampup), self.consistency_rampup)

    _SLOT = dict(t1=0, t2=1, t3=2, s0=0, lb=1, ul=2, lu=3, s=4, lq=5)      # forward order per model (train.py:643-647,668,699-702,740)

    def _multi(self):
        return self.lanes > 1 and self._eager_steps > 0 and E.PROFILE_EVENTS is None

    def _run_jobs(self, jobs):
        """Enqueue independent jobs round-robin on the lane streams (fork from / join into the current stream)."""
        K = min(self.lanes, len(jobs)) if self._multi() else 1
        if K <= 1:
            return [j() for j in jobs]
        main = torch.cuda.current_stream()
        while len(self._lane_streams) < K:
            self._lane_streams.append(torch.cuda.Stream())
        ev = torch.cuda.Event()
        ev.record(main)
        results = []
        for i, job in enumerate(jobs):
            st = self._lane_streams[i % K]
            if i < K:
                st.wait_event(ev)
            with torch.cuda.stream(st):
                results.append(job())
        for st in self._lane_streams[:K]:
            main.wait_stream(st)
        return results

    def _forward(self, model, a: E.Act, need_grad: bool, tag: str = None):
        multi = self._multi()
        ctx = E.Ctx(model.training, need_grad, bn_sync=bridge.BN_SYNC, stats=self._stats if multi else None, slot=self._SLOT.get(tag, 0),
                    grads_on_side=multi)
        ctx.bn_world = bridge.BN_WORLD
        kw = self.forward_kwargs
        if self._domains is not None:
            kw = dict(kw, domain_label=self._domains[tag])
        (logits, head_bwd), = model.program(ctx, a, **kw)
        return ctx, logits, head_bwd

    def _grad_provider(self, param):
        self._touched.add(id(param))
        return self.opt.grad_for(param)

    def _grad_done(self, param):
        if self.dp is not None:
            self.dp.on_grad_done(param, self._last_branch)

    def _branch(self, a, target_u8, mask_u8, weight_idx, tag):
        """forward -> loss -> backward of one loss branch; returns loss3 device tensor.  ``weight_idx``: slot of the hyper
        block holding this term's outer weight (1, cw, cw, cw^2: train.py:838)."""
        ctx, logits, head_bwd = self._forward(self.model, a, True, tag)
        if self.dp is not None and self.dp.global_loss and self.dp.active:
            loss3, coef = term_forward(logits, target_u8, mask_u8, self.branch, allreduce=self.dp.sum_across_ranks, world=self.dp.world)
        else:
            loss3, coef = term_forward(logits, target_u8, mask_u8, self.branch)
        dlogits = term_backward(logits, target_u8, mask_u8, self.branch, coef, upstream=self.hyper[weight_idx:weight_idx + 1])
        sink = E.GradSink(provider=self._grad_provider, done=self._grad_done)
        head_bwd(dlogits, sink)
        ctx.backward(sink, join=not self._multi())       # multi-lane: one join before the optimiser
        if self.dp is not None:
            self.dp.on_branch_done()
        return loss3, logits

    def _last_branch_job(self, a, target_u8, mask_u8):
        self._last_branch = True          # from here on finished gradient buckets may be all-reduced
        try:
            return self._branch(a, target_u8, mask_u8, H_CW2, "s")
        finally:
            self._last_branch = False

    def _set_hyper(self, it, gscale):
        """Write this step's scalars (host double arithmetic exactly as train.py:819-820,838,91,854) to the device block."""
        cw = self.consistency_weight(it)
        alpha = min(1 - 1 / (it + 1), self.ema_decay)
        slot = self._hyper_slot
        self._hyper_slot = (slot + 1) % _HYPER_RING
        if self._hyper_ev[slot] is not None:
            self._hyper_ev[slot].synchronize()          # the copy that last used this pinned slot (4 steps ago) has completed
        h = self._hyper_host[slot]
        h[H_LR], h[H_ALPHA], h[H_GSCALE], h[H_CW], h[H_CW2], h[H_ONE] = self.lr, alpha, gscale, cw, cw * cw, 1.0
        self.hyper.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_ev[slot] = ev
        return cw

    # -- input prefetch ------------------------------------------------------------------------------------------------
    def upload(self, batch, lq=None):
        """Start the host-to-device copy of a (pinned) host batch on the trainer's copy stream and return the batch to pass
        to ``step``: called right after ``step(i)`` was issued, the upload of step i+1 overlaps step i's kernels (the reference
        copies synchronously, train.py:582-589).  Two staging sets alternate; ``step`` waits for the copy's event, not for the
        host."""
        dev = self.hyper.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        slot = self._up_slot
        self._up_slot ^= 1
        items = dict(batch)
        if lq is not None:
            items["_lq_plain"] = lq
        bufs = self._staging[slot]
        fresh = False
        for k, v in items.items():
            if isinstance(v, torch.Tensor) and (k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype):
                bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                fresh = True
        st = self._copy_stream
        if fresh:
            # a new staging buffer may be a block the allocator has just taken back from kernels of the CURRENT stream that are
            # still in flight (eager steps): the copy stream must not write it before they are done
            st.wait_stream(torch.cuda.current_stream())
        if self._consumed[slot] is not None:
            st.wait_event(self._consumed[slot])                 # the step that read this staging set has been issued and must finish first
        out = {}
        with torch.cuda.stream(st):
            for k, v in items.items():
                if isinstance(v, torch.Tensor):
                    bufs[k].copy_(v, non_blocking=True)
                    out[k] = bufs[k]
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(st)
        out["_ustrun_ready"] = (ev, slot)
        return out

    # -- checkpointing (utils/util.py:259-297 saves the optimizer state; train.py:542-548 restores it and iter_num) ----------
    def state_dict(self):
        return {"optimizer": self.opt.state_dict(), "iter_num": self.iter_num, "lr": self.lr}

    def load_state_dict(self, sd):
        self.opt.load_state_dict(sd["optimizer"])
        self.iter_num, self.lr = int(sd["iter_num"]), float(sd["lr"])
        self._graphs = {}

    # -- the step ----------------------------------------------------------------------------
    def step(self, batch, lq=None, keep_logits=False):
        """batch: lb_x [Bl,C,H,W] fp32, lb_mask ([Bl,H,W] int | [Bl,C,H,W] float), ulb_w, ulb_s,
        move_transx [Bu,C,H,W] (or, instead, mix_ratio [Bu]: the style mix then runs on the device),
        box [Bu,H,W] {0,1}, choice [Bu] int, cut_img [Nc,C,H,W],
        cut_label, cut_mask (labelled batch + confidence bank).  ``lq``: optional [1,C,H,W] image
        whose student forward only updates BN running statistics (train.py:740).
        DSBN networks: ``domain_lb`` / ``domain_ulb`` (host ints; see ``step_domains``).
        Returns device tensors (losses, compositions); nothing is copied to the host.  With ``use_graph`` the returned
        tensors are the graph's static outputs: read them before the next ``step``."""
        it = self.iter_num
        ready = batch.get("_ustrun_ready")
        if ready is not None:                       # inputs staged by upload(): wait for the copy on the device, not on the host
            torch.cuda.current_stream().wait_event(ready[0])
            if lq is None:
                lq = batch.get("_lq_plain")
            batch = {k: v for k, v in batch.items() if k not in ("_ustrun_ready", "_lq_plain")}
        if self.dp is not None and getattr(self.dp, "peer", None) is not None:
            self.dp.peer.poll()                     # a peer-BN timeout of an earlier step raises here (one step late, no sync)
            self.dp.peer.begin_step()               # device-side sequence base of this step's cross-rank BN barriers
        gscale = 1.0 if self.dp is None else (1.0 if self.dp.global_loss else 1.0 / self.dp.world)
        domains = None
        if "domain_lb" in batch or "domain_ulb" in batch:
            domains = step_domains(int(batch["domain_lb"]), int(batch["domain_ulb"]))
        tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
        if lq is not None and not isinstance(lq, torch.Tensor):       # (lq_u, labelled image, box): ConfidenceBank.lq_input
            tensors["_lq_u"], tensors["_lq_img"], tensors["_lq_box"] = lq[0], lq[1], as_u8(lq[2])
            lq = None
        if "mix_ratio" in batch and "mix_ratio" not in tensors:       # host floats (train.py:180) -> one float64 per sample
            tensors["mix_ratio"] = torch.as_tensor(batch["mix_ratio"], dtype=torch.float64)
        key = None
        if self.use_graph:
            dom_key = None if domains is None else tuple(domains[t][0] for t in STEP_DOMAIN_TAGS)
            key = (tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in tensors.items())), None if lq is None else tuple(lq.shape), bool(keep_logits),
                   dom_key, bool(self.first_epoch), self.hardness_mode, E.get_precision())
        # the first two steps of every input signature run eagerly: they size the allocator pool, initialise the momentum
        # buffers (the optimiser's descriptor table changes after step 0) and set the per-kernel launch attributes
        if key is not None and self._eager_by_key.get(key, 0) >= 2:
            out, cw = self._step_graphed(key, tensors, lq, keep_logits, domains, it, gscale)
        else:
            cw = self._set_hyper(it, gscale)
            k0 = E.KERNELS
            out = self._step_body(tensors, lq, keep_logits, domains)
            self.launches_per_step = E.KERNELS - k0
            self._eager_steps += 1
            if key is not None:
                self._eager_by_key[key] = self._eager_by_key.get(key, 0) + 1
        if ready is not None:
            ev = torch.cuda.Event()
            ev.record()
            self._consumed[ready[1]] = ev
        # the reference's lr schedule (train.py:854-858)
        self.lr = self.base_lr * (1.0 - it / self.max_iterations) ** 0.9
        self.iter_num = it + 1
        out = dict(out)
        out["consistency_weight"] = cw
        return out

    def _step_graphed(self, key, tensors, lq, keep_logits, domains, it, gscale):
        """Replay the step as ONE CUDA graph (captured on first use per input signature): ~1300 (UNet-A) / ~2000 (UNet-B)
        kernel launches leave the host as a single cudaGraphLaunch; inputs are copied into the capture's static buffers."""
        if self.dp is not None and not self.dp.graph_safe:
            raise NotImplementedError("CUDA-graph replay of this data-parallel configuration is not enabled")
        ent = self._graphs.get(key)
        if ent is None:
            static = {k: v.detach().to(self.hyper.device, copy=True) for k, v in tensors.items()}
            static_lq = None if lq is None else lq.detach().to(self.hyper.device, copy=True)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()                 # the eager warm-up's cached blocks go back to the driver: the capture gets its own pool
            g = torch.cuda.CUDAGraph()
            k0 = E.KERNELS
            peer = getattr(self.dp, "peer", None) if self.dp is not None else None
            try:
                with torch.cuda.graph(g):
                    out = self._step_body(static, static_lq, keep_logits, domains)
            except Exception as e:
                if self.dp is None:
                    raise
                # data parallel: the capture includes NCCL calls; if this platform cannot capture them, say so and keep stepping
                # eagerly (an eager rank issues the same collectives in the same order as a replaying one)
                import warnings
                warnings.warn(f"CUDA-graph capture of the data-parallel step failed ({type(e).__name__}: {e}); continuing without graphs")
                self.use_graph, self.graph_error = False, f"{type(e).__name__}: {e}"
                if peer is not None:
                    peer.calls = 0
                cw = self._set_hyper(it, gscale)
                return self._step_body(tensors, lq, keep_logits, domains), cw
            peer = getattr(self.dp, "peer", None) if self.dp is not None else None
            ent = self._graphs[key] = (g, static, static_lq, out, E.KERNELS - k0, list(self.opt.has_grad), 0 if peer is None else peer.calls)
        g, static, static_lq, out, self.launches_per_step, has_grad, peer_calls = ent
        if self.dp is not None and getattr(self.dp, "peer", None) is not None:
            self.dp.peer.calls = peer_calls          # cross-rank BN barriers this replay executes (begin_step folds them into the base)
        for i, f in enumerate(has_grad):             # host-side bookkeeping of the optimiser follows the graph that runs
            self.opt.set_has_grad(i, f)
            if f:
                self.opt.first[i] = False
        for k, v in tensors.items():
            static[k].copy_(v, non_blocking=True)
        if lq is not None:
            static_lq.copy_(lq, non_blocking=True)
        cw = self._set_hyper(it, gscale)
        g.replay()
        return out, cw

    def _step_body(self, b, lq, keep_logits, domains):
        """Device work of one step on the current stream; every per-step scalar comes from ``self.hyper``."""
        self._domains = domains
        branch = self.branch
        dev = b["ulb_w"].device
        box_u8 = as_u8(b["box"]).contiguous()
        inv_box = 1 - box_u8
        choice_i = b["choice"].to(device=dev, dtype=torch.int32).contiguous()
        if "move_transx" not in b:
            # frequency-domain style mix on the device (train.py:628-636): the host only supplies
            # its random blend ratio per sample (train.py:180) instead of running fft2/ifft2 in numpy
            from .fft_mix import amp_mix
            b = dict(b)
            cut_f = normalize_u8(b["cut_img"]) if _is_u8_image(b["cut_img"]) else b["cut_img"].float()
            ulb_f = normalize_u8(b["ulb_w"]) if _is_u8_image(b["ulb_w"]) else b["ulb_w"]
            b["move_transx"] = amp_mix(cut_f[choice_i.long()], ulb_f, b["mix_ratio"], self.fft_window)
        multi = self._multi()
        if self._eager_steps > 0:
            self._pack_plan.run()                     # every packed weight copy in one launch, on this stream, before the lanes fork
        if multi:
            self._stats.begin_step()
        # 1. teacher (train-mode BN, no grad) and 2. student on ulb_w (no grad): four independent forwards
        fwd = lambda model, tag, *mix: (lambda: self._forward(model, mix_input(*mix), False, tag)[1])
        t1, t2, t3, s0 = self._run_jobs([fwd(self.ema_model, "t1", b["ulb_w"], None, None),
                                          fwd(self.ema_model, "t2", b["ulb_w"], b["cut_img"], box_u8, choice_i),
                                          fwd(self.ema_model, "t3", b["ulb_w"], b["cut_img"], inv_box, choice_i),
                                          fwd(self.model, "s0", b["ulb_w"], None, None)])
        # 3. pseudo labels
        comp = pseudo_labels(t1, t2, t3, box_u8, b["cut_label"], b["cut_mask"], choice_i, self.threshold, branch, student_logits=s0)
        # 4. four loss branches (+ the batch-1 low-quality forward, which only feeds BatchNorm running statistics)
        self.opt.zero_grad()
        self._touched = set()
        if self.dp is not None:
            self.dp.begin_step(self.opt)
        lb_t = as_u8(b["lb_mask"]).contiguous()
        jobs = [lambda: self._branch(mix_input(b["lb_x"], None, None), lb_t, None, H_ONE, "lb"),
                lambda: self._branch(mix_input(b["ulb_s"], b["move_transx"], box_u8), comp["pseudo_label_ul"], comp["mask_ul"], H_CW, "ul"),
                lambda: self._branch(mix_input(b["move_transx"], b["ulb_s"], box_u8), comp["pseudo_label_lu"], comp["mask_lu"], H_CW, "lu"),
                lambda: self._last_branch_job(mix_input(b["ulb_s"], None, None), comp["pseudo_label_w"], comp["mask_w"])]
        if lq is not None or "_lq_u" in b:
            # lq: the image itself, or (lq_u, labelled image, box) = the CutMix of train.py:731 composed by the input kernel
            lq_mix = (lq, None, None) if lq is not None else (b["_lq_u"], b["_lq_img"], b["_lq_box"].contiguous())
            jobs.append(lambda: self._forward(self.model, mix_input(*lq_mix), False, "lq")[1])
        res = self._run_jobs(jobs)
        (l_sup, lg_lb), (l_ul, lg_ul), (l_lu, lg_lu), (l_s, lg_s) = res[:4]
        if multi:
            E.join_side_stream()                      # every weight / BatchNorm-parameter gradient is complete
            self._stats.flush()                       # running statistics in the reference's forward order
        # 5. data-parallel gradient reduction (averaging folded into the optimiser's grad_scale)
        if self.dp is not None:
            self.dp.finish_step(self.opt)
        for i, p in enumerate(self.params):
            self.opt.set_has_grad(i, id(p) in self._touched)
        # 6. fused SGD + EMA (lr, alpha, gradient scale from the hyper block)
        self.opt.step(hyper=self.hyper)
        cw_t = self.hyper[H_CW]
        loss = l_sup[0] + cw_t * (l_ul[0] + l_lu[0] + cw_t * l_s[0])
        out = dict(comp)
        if self.hardness_mode is not None:
            out["hardness"], out["lq_idx"], out["stu_tea_dice"] = hardness(comp["stu_pseudo_label"], comp["pseudo_label"], self.hardness_mode, self.first_epoch)
        out.update(loss=loss, sup_loss=l_sup[0], unsup_loss_ul=l_ul[0], unsup_loss_lu=l_lu[0], unsup_loss_s=l_s[0])
        if keep_logits:
            out["logits"] = dict(t1=t1, t2=t2, t3=t3, s0=s0, lb=lg_lb, ul=lg_ul, lu=lg_lu, s=lg_s)
        self._domains = None
        return out


HARDNESS_MODES = {"binary": 0, "2label": 1, "3label": 2}


def hardness(stu_pseudo_label: torch.Tensor, pseudo_label: torch.Tensor, mode: str = "binary", first_epoch: bool = False):
    """Per-sample hardness of the unlabelled batch on the device (train.py:705-718): ``1 - Dice(student pseudo label,
    teacher pseudo label)`` with the reference's ``dice_coefficient_numpy`` formula (utils/metrics.py:114-146), averaged
    over the dataset's label parts, plus the index of the hardest ("low quality") sample.  ``mode``: "binary"
    (metrics.dice_coeff: prostate / BUSI), "2label" (dice_coeff_2label: fundus, planes [B,2,H,W]), "3label"
    (dice_coeff_3label: M&Ms).  Returns (hardness float64 [B], lq_idx int32 [1], dice float64 [parts,B]) -- device
    tensors, no host synchronisation (the reference copies both label maps to the host here)."""
    L.require_device()
    m = HARDNESS_MODES[mode]
    s, t = as_u8(stu_pseudo_label).contiguous(), as_u8(pseudo_label).contiguous()
    if s.shape != t.shape or s.dim() != (4 if m == 1 else 3):
        raise ValueError("hardness: label planes must have the same shape ([B,H,W]; [B,2,H,W] for '2label')")
    if m == 1 and s.shape[1] != 2:
        raise ValueError("hardness: '2label' expects two channels")
    B, H, W = s.shape[0], s.shape[-2], s.shape[-1]
    parts = (1, 2, 3)[m]
    dev = s.device
    ws = torch.empty(B * 9, dtype=torch.int32, device=dev)
    hard = torch.empty(B, dtype=torch.float64, device=dev)
    dice = torch.empty((parts, B), dtype=torch.float64, device=dev)
    lq = torch.empty(1, dtype=torch.int32, device=dev)
    _call("ustrun_hardness", _ptr(s), _ptr(t), B, H, W, m, 1 if first_epoch else 0, _ptr(ws), _ptr(hard), _ptr(dice), _ptr(lq), _stream())
    return hard, lq, dice
