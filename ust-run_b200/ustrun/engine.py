"""Host-side execution engine: NHWC activations, layer ops and a layer-level backward tape.

The nn.Module drop-ins (networks/*.py) describe a network as a sequence of these ops; every op
launches kernels of libustrun_sm100.so through the C ABI (``_lib.call``) on torch's current CUDA
stream.  PyTorch is used for memory (torch.empty), streams and parameter storage only.

Precision modes (``set_precision``):
  * ``"bf16"`` (default): bf16 NHWC activations, fp32 accumulation; convs with Cin,Cout % 64 == 0
    run on the tcgen05 kernels, the narrow ones (first conv, logits head, 16/32-channel UNet-B
    levels) on the CUDA-core kernels.
  * ``"fp32"``: validation mode, everything fp32 on the CUDA-core kernels (1e-4 parity target).
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Callable, Dict, List, Optional

import torch

from . import _lib as L

_PRECISION = os.environ.get("USTRUN_PRECISION", "bf16")
# explicit = chosen by USTRUN_PRECISION or set_precision().  Otherwise the Tier-A bridge follows the caller's autocast state
# (train.py --amp 1: autocast -> bf16 storage; --amp 0: no autocast -> the fp32 validation mode), the fused step uses bf16.
_PRECISION_EXPLICIT = "USTRUN_PRECISION" in os.environ
_FORCE_SIMT = os.environ.get("USTRUN_FORCE_SIMT", "0") == "1"
LAUNCHES = 0          # C-ABI calls issued
KERNELS = 0           # kernels those calls launched (bench.py reports it as gpu_launches)
PROFILE_EVENTS = None  # bench.py sets this to a list: (kernel class, algorithmic FLOPs, start event, end event)
MANIFEST = None        # tools/step_once.py sets this to a list: one dict per profiled C-ABI call, in launch order (for ncu summaries)
_KERNELS_PER_CALL = {"ustrun_conv_wgrad": 2, "ustrun_convT2x2_wgrad": 2, "ustrun_convT2x2_fwd": 1, "ustrun_channel_sum": 2,
                     "ustrun_ce_dice_softmax_fwd": 2, "ustrun_bce_dice_sigmoid_fwd": 2}


def set_precision(mode: str) -> None:
    global _PRECISION, _PRECISION_EXPLICIT
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = mode
    _PRECISION_EXPLICIT = True


class precision_scope:
    """Temporarily run engine ops in ``mode`` (None: leave as is)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global _PRECISION
        self.prev = _PRECISION
        if self.mode is not None:
            _PRECISION = self.mode

    def __exit__(self, *exc):
        global _PRECISION
        _PRECISION = self.prev


def tier_a_precision() -> str:
    """Precision of a module call through the autograd bridge (Tier A)."""
    if _PRECISION_EXPLICIT:
        return _PRECISION
    return "bf16" if torch.is_autocast_enabled("cuda") else "fp32"


def get_precision() -> str:
    return _PRECISION


def set_force_simt(flag: bool) -> None:
    global _FORCE_SIMT
    _FORCE_SIMT = bool(flag)


def _dt():
    return (torch.bfloat16, L.BF16) if _PRECISION == "bf16" else (torch.float32, L.F32)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream as an int (ctypes converts it for the c_void_p parameter); the raw
    accessor avoids building a torch.cuda.Stream object on each of the ~1300 calls of a step."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device()) or None
    return torch.cuda.current_stream().cuda_stream or None


def _call(name, *args):
    global LAUNCHES, KERNELS
    LAUNCHES += 1
    KERNELS += _KERNELS_PER_CALL.get(name, 1)
    L.call(name, *args)


def _profiled(cls, flops, name, *args, meta=None):
    """_call bracketed by CUDA events on the launching stream when bench.py profiles.  ``flops``: algorithmic FLOPs (tensor
    classes) or algorithmic bytes (``hbm_*`` classes) of the call."""
    if MANIFEST is not None:
        MANIFEST.append({"cls": cls, "entry": name, "work": float(flops), "kernels": _KERNELS_PER_CALL.get(name, 1), "meta": meta})
    if PROFILE_EVENTS is None:
        return _call(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _call(name, *args)
    e1.record()
    PROFILE_EVENTS.append((cls, flops, e0, e1))


# Weight-gradient GEMMs are off the backward critical path (nothing downstream reads dW until the
# optimiser / all-reduce) and use ~56 registers per thread, so they run on a side stream and share the
# SMs with the HBM-bound BatchNorm / pooling kernels of the next layers instead of serialising with them.
_WGRAD_OVERLAP = os.environ.get("USTRUN_WGRAD_OVERLAP", "1") == "1"
_WGRAD_AFTER_DGRAD = os.environ.get("USTRUN_WGRAD_AFTER_DGRAD", "0") == "1"
_SIDE: Dict[int, "torch.cuda.Stream"] = {}
_SIDE_MAX_BYTES = int(os.environ.get("USTRUN_WGRAD_OVERLAP_MAX_BYTES", str(768 << 20)))


def set_wgrad_overlap(flag: bool) -> None:
    global _WGRAD_OVERLAP
    _WGRAD_OVERLAP = bool(flag)


def _side_stream():
    if not _WGRAD_OVERLAP or PROFILE_EVENTS is not None:
        return None
    dev = torch.cuda.current_device()
    st = _SIDE.get(dev)
    if st is None:
        st = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return st


_SYNC: Dict[int, "torch.cuda.Stream"] = {}


def on_sync_stream(fn: Callable) -> None:
    """Run ``fn`` (a cross-rank BatchNorm finalize kernel that WAITS for the other ranks) on the one BN-sync stream, ordered
    after the current stream's work, and make the current stream wait for it.  In the multi-lane data-parallel step every
    such kernel of every lane goes through this stream in program order, so all ranks execute them in the same order whatever
    the lanes do -- two ranks can never wait for each other's kernels in crossed order."""
    dev = torch.cuda.current_device()
    st = _SYNC.get(dev)
    if st is None:
        st = _SYNC[dev] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    with torch.cuda.stream(st):
        st.wait_event(ev)
        fn()
        done = torch.cuda.Event()
        done.record(st)
    cur.wait_event(done)


_HELD: List = []          # operands of side-stream kernels enqueued during a CUDA-graph capture (see on_side_stream)
_SIDE_DIRTY = [False]     # the side stream has work the current stream has not joined yet


def on_side_stream(fn: Callable, tensors) -> None:
    """Run ``fn`` (which enqueues kernels reading ``tensors``) on the weight-gradient stream, ordered after
    everything enqueued so far on the current stream."""
    side = _side_stream()
    # record_stream keeps the operands out of the allocator until the side stream has caught up: for multi-GB activations
    # (cfg5: 2 GB per layer-1 tensor) that inflates the footprint by tens of GB for no gain, so those run in line
    if side is None or sum(t.numel() * t.element_size() for t in tensors) > _SIDE_MAX_BYTES:
        fn()
        return
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(side):
        side.wait_event(ev)               # during a capture this forks the side stream into the graph
        fn()
    _SIDE_DIRTY[0] = True
    if torch.cuda.is_current_stream_capturing():
        # record_stream would keep the blocks out of the capture's private pool until the capture ends (its events cannot be
        # queried inside a capture): hold references until the join instead -- allocations after the join are ordered after it
        _HELD.extend(tensors)
    else:
        for t in tensors:
            t.record_stream(side)           # the caching allocator must not recycle them under the side stream


def join_side_stream() -> None:
    """Make the current stream wait for all weight-gradient kernels (before the optimiser / all-reduce)."""
    st = _SIDE.get(torch.cuda.current_device()) if torch.cuda.is_available() else None
    if st is not None and _SIDE_DIRTY[0]:
        torch.cuda.current_stream().wait_stream(st)
        _SIDE_DIRTY[0] = False
    _HELD.clear()


def reserve_pool(nbytes: Optional[int] = None, fraction: float = 0.5, cap: int = 96 << 30) -> int:
    """Seed torch's caching allocator with one large free block (default: half of the free HBM, at most 96 GiB).
    The engine allocates every activation through torch; without this the pool grows by cudaMalloc (a device
    synchronisation each) over the first ~20 steps -- longer with the weight-gradient side stream, whose
    ``record_stream`` holds blocks back -- and those steps run 2-3x slower.  Returns the bytes reserved."""
    L.require_device()
    if nbytes is None:
        free, _total = torch.cuda.mem_get_info()
        nbytes = min(int(free * fraction), cap)
    nbytes = max(int(nbytes), 0)
    if nbytes:
        blk = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        del blk                      # stays cached; later requests split it
    return nbytes


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class Act:
    """An NHWC activation: ``t`` is a contiguous [B,H,W,ld] tensor, the activation is channels
    [c0, c0+C) of it (so an Act may be one half of a concat buffer)."""

    __slots__ = ("t", "B", "H", "W", "C", "c0", "parent", "_g", "needs_grad")

    def __init__(self, t, C=None, c0=0, parent=None):
        self.t = t
        self.B, self.H, self.W = t.shape[0], t.shape[1], t.shape[2]
        self.C = t.shape[3] if C is None else C
        self.c0 = c0
        self.parent = parent
        self._g = None
        self.needs_grad = False

    @staticmethod
    def new(B, H, W, C, dtype=None, device=None):
        return Act(torch.empty((B, H, W, C), dtype=dtype or _dt()[0], device=device or "cuda"))

    def like(self, C=None):
        return Act(torch.empty((self.B, self.H, self.W, C or self.C), dtype=self.t.dtype, device=self.t.device))

    def view(self, c0, C):
        return Act(self.t, C, self.c0 + c0, parent=self)

    @property
    def ld(self):
        return self.t.shape[3]

    @property
    def npix(self):
        return self.B * self.H * self.W

    @property
    def ptr(self):
        return self.t.data_ptr() + self.c0 * self.t.element_size()

    @property
    def dtype_code(self):
        return L.BF16 if self.t.dtype == torch.bfloat16 else L.F32

    # gradient buffer: views resolve through their parent (concat halves)
    @property
    def g(self):
        if self.parent is not None:
            pg = self.parent.g
            return None if pg is None else Act(pg.t, self.C, pg.c0 + (self.c0 - self.parent.c0), parent=None)
        return self._g

    @g.setter
    def g(self, value):
        if self.parent is not None:
            raise RuntimeError("cannot assign the gradient of a view")
        self._g = value


class GradSink:
    """Where weight gradients go.  ``get(param)`` returns (fp32 tensor, accumulate_flag)."""

    def __init__(self, provider: Optional[Callable] = None, done: Optional[Callable] = None):
        self.provider = provider
        self._done = done
        self.fresh: Dict[int, torch.Tensor] = {}

    def done(self, *params):
        """The kernel that writes these parameters' gradients has just been enqueued on the CURRENT stream (data-parallel
        gradient buckets may be all-reduced from here on, ordered after that stream)."""
        if self._done is not None:
            for p in params:
                if p is not None:
                    self._done(p)

    def get(self, param: torch.Tensor):
        if self.provider is not None:
            return self.provider(param), 1
        key = id(param)
        if key in self.fresh:
            return self.fresh[key], 1
        g = torch.empty_like(param, dtype=torch.float32)
        self.fresh[key] = g
        return g, 0


class Ctx:
    """One forward pass: holds the backward tape when gradients are needed."""

    def __init__(self, training: bool, need_grad: bool, bn_sync: Optional[Callable] = None, stats: Optional["StatLog"] = None,
                 slot: int = 0, grads_on_side: bool = False):
        self.training = training
        self.need_grad = need_grad
        self.tape: List[Callable] = []
        self.bn_sync = bn_sync           # callable(tensor[2*C]) -> None (in-place cross-rank sum) or None
        self.bn_world = 1
        self.bn_peer = getattr(bn_sync, "peer", None)   # PeerStats: fused finalize + NVLink peer-memory reduction
        # multi-lane step: BatchNorm running statistics are logged per forward (slot) and applied in the reference's order at
        # the end of the step; every parameter-gradient accumulation goes to the ONE weight-gradient stream in program order
        self.stats = stats
        self.slot = slot
        self.grads_on_side = grads_on_side

    def backward(self, sink: GradSink, join: bool = True):
        """``join=False`` leaves the weight-gradient side stream running (the fused step joins once, before the
        optimiser, so the tail of one branch's weight gradients overlaps the next branch's forward)."""
        for fn in reversed(self.tape):
            fn(sink)
        self.tape = []
        if join:
            join_side_stream()           # weight gradients are complete for whoever runs next on this stream


class StatLog:
    """Deferred BatchNorm running statistics of one step (see ustrun_bn_running_update): ``slot_ptr`` hands the finalize
    kernel the slot of (layer, forward); ``flush`` applies every layer's logged forwards in slot order with one launch."""
    SLOTS = 8

    def __init__(self):
        self.entries: Dict[int, list] = {}        # id(bn) -> [bn, stats tensor [SLOTS][2][C], mask]
        self._tables: Dict[tuple, tuple] = {}

    def begin_step(self):
        for e in self.entries.values():
            e[2] = 0

    def slot_ptr(self, bn, slot: int) -> int:
        e = self.entries.get(id(bn))
        if e is None:
            C = bn.num_features
            e = self.entries[id(bn)] = [bn, torch.zeros((self.SLOTS, 2, C), dtype=torch.float32, device=bn.running_mean.device), 0]
        e[2] |= 1 << slot
        return e[1].data_ptr() + slot * 2 * bn.num_features * 4

    def flush(self):
        import numpy as np
        live = [e for e in self.entries.values() if e[2]]
        if not live:
            return
        key = tuple((id(e[0]), e[2], e[0].running_mean.data_ptr()) for e in live)
        ent = self._tables.get(key)
        if ent is None:
            dt = np.dtype([("rm", "<u8"), ("rv", "<u8"), ("nbt", "<u8"), ("stats", "<u8"), ("C", "<i4"), ("mask", "<u4"), ("mom", "<f4"), ("pad", "<i4")])
            tab = np.zeros(len(live), dtype=dt)
            for i, (bn, stats, mask) in enumerate(live):
                tab[i]["rm"], tab[i]["rv"] = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                tab[i]["nbt"] = bn.num_batches_tracked.data_ptr() if bn.num_batches_tracked is not None else 0
                tab[i]["stats"], tab[i]["C"], tab[i]["mask"] = stats.data_ptr(), bn.num_features, mask
                tab[i]["mom"] = 0.1 if bn.momentum is None else float(bn.momentum)
            host = torch.from_numpy(tab.view(np.uint8).reshape(-1).copy()).pin_memory()     # pinned + stream-ordered: legal in a capture
            t = torch.empty(host.numel(), dtype=torch.uint8, device=live[0][1].device)
            t.copy_(host, non_blocking=True)
            ent = self._tables[key] = (t, len(live), max(e[0].num_features for e in live), host)
        t, n, cmax = ent[:3]
        _call("ustrun_bn_running_update", _ptr(t), n, cmax, _stream())


def collect_packed(*models):
    """Every PackedConv of the given modules (attributes named ``_pk*``: one object or a list)."""
    out = []
    for model in models:
        for m in model.modules():
            for name, v in vars(m).items():
                if name.startswith("_pk"):
                    out.extend(v if isinstance(v, (list, tuple)) else [v])
    return [p for p in out if isinstance(p, PackedConv)]


class PackPlan:
    """All packed weight copies of a set of convs refreshed by ONE launch (ustrun_pack_weights_multi) on the current stream:
    after the optimiser step every copy of the student and of the teacher is stale, and packing them lazily is 46 (UNet-A) to
    ~200 (UNet-B) few-microsecond launches per step.  Only convs that have run before are known (``PackedConv.last``); the
    others keep packing lazily on first use."""

    def __init__(self, packed_list):
        self.packed = packed_list
        self._sig = None
        self._dev = None

    @staticmethod
    def _key(weight, code):
        return (weight.data_ptr(), weight._version, code, getattr(weight, "_ustrun_epoch", 0))

    def run(self):
        import numpy as np
        tdt, code = _dt()
        ents = []
        for pk in self.packed:
            if pk.last is None:
                continue
            w = pk.last[0]()
            if w is None or not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous():
                continue
            ents.append((pk, w, pk.last[1], pk.last[2]))
        if not ents or all(pk.key == self._key(w, code) and (pk.has_wd or not nwd) for pk, w, _, nwd in ents):
            return
        sig = tuple((id(pk), w.data_ptr(), code, tr, nwd) for pk, w, tr, nwd in ents)
        if sig != self._sig:
            dt = np.dtype([("w", "<u8"), ("wf", "<u8"), ("wd", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"), ("taps", "<i4"), ("tr", "<i4")])
            tab = np.zeros(len(ents), dtype=dt)
            be, bt = [], []
            for i, (pk, w, tr, nwd) in enumerate(ents):
                if tr:
                    cin, cout, taps = w.shape[0], w.shape[1], 4
                    fshape, dshape = (4, cout, cin), (cin, 4, cout)
                    ntiles = (cin * cout * 4 + 4095) // 4096
                else:
                    cout, cin, taps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
                    fshape, dshape = (cout, taps, cin), (cin, taps, cout)
                    ntiles = ((cout + 31) // 32) * ((cin + 31) // 32)
                if pk.wf is None or pk.wf.dtype != tdt or tuple(pk.wf.shape) != fshape:
                    pk.wf = torch.empty(fshape, dtype=tdt, device=w.device)
                if nwd and (pk.wd is None or pk.wd.dtype != tdt or tuple(pk.wd.shape) != dshape):
                    pk.wd = torch.empty(dshape, dtype=tdt, device=w.device)
                tab[i] = (w.data_ptr(), pk.wf.data_ptr(), pk.wd.data_ptr() if nwd else 0, cout, cin, taps, 1 if tr else 0)
                be += [i] * ntiles
                bt += list(range(ntiles))
            dev = ents[0][1].device
            up = lambda a: torch.from_numpy(a).pin_memory()
            host = (up(tab.view(np.uint8).reshape(-1).copy()), up(np.asarray(be, np.int32)), up(np.asarray(bt, np.int32)))
            self._dev = tuple(torch.empty_like(h, device=dev) for h in host) + (len(be), host)
            for d, h in zip(self._dev[:3], host):
                d.copy_(h, non_blocking=True)
            self._sig = sig
        t, e, b, n, _ = self._dev
        _call("ustrun_pack_weights_multi", _ptr(t), _ptr(e), _ptr(b), n, code, _stream())
        for pk, w, _, nwd in ents:
            pk.key = self._key(w, code)
            pk.has_wd = bool(nwd)


def prepack(packed_list):
    """Refresh the packed weight copies on the CURRENT stream (multi-lane step: before the lanes fork, so that no lane
    depends on a packing kernel another lane launched).  Only convs that have run before are known."""
    for pk in packed_list:
        if pk.last is not None:
            w = pk.last[0]()
            if w is not None:
                pk.get(w, transposed=pk.last[1], need_wd=pk.last[2])


# --------------------------------------------------------------------------------------------
# packed weights (bf16/fp32 [Cout][tap][Cin] forward and [Cin][tap'][Cout] dgrad copies)
# --------------------------------------------------------------------------------------------
class PackedConv:
    def __init__(self):
        self.key = None
        self.wf = None
        self.wd = None
        self.has_wd = False
        self.last = None          # (weakref to the weight, transposed, need_wd) of the last get(): what prepack() replays

    def get(self, weight: torch.Tensor, transposed=False, need_wd=True):
        """(wf, wd) for ``weight``.  ``need_wd=False`` (no-grad forwards: the teacher never runs a dgrad) packs only the
        forward copy; the dgrad copy is added lazily the first time a caller asks for it."""
        if self.last is None or self.last[0]() is not weight or (need_wd and not self.last[2]):
            self.last = (weakref.ref(weight), transposed, bool(need_wd))
        tdt, code = _dt()
        key = (weight.data_ptr(), weight._version, code, getattr(weight, "_ustrun_epoch", 0))
        fresh = key != self.key
        if fresh or (need_wd and not self.has_wd):
            w = weight.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            if transposed:
                cin, cout = w.shape[0], w.shape[1]
                if fresh:
                    self.wf = torch.empty((4, cout, cin), dtype=tdt, device=w.device)
                self.wd = torch.empty((cin, 4, cout), dtype=tdt, device=w.device) if need_wd else None
                _call("ustrun_pack_convT_weight", _ptr(w), _ptr(self.wf) if fresh else None, _ptr(self.wd), code, cin, cout, _stream())
            else:
                cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
                if fresh:
                    self.wf = torch.empty((cout, k * k, cin), dtype=tdt, device=w.device)
                self.wd = torch.empty((cin, k * k, cout), dtype=tdt, device=w.device) if need_wd else None
                _call("ustrun_pack_conv_weight", _ptr(w), _ptr(self.wf) if fresh else None, _ptr(self.wd), code, cout, cin, k, _stream())
            self.key = key
            self.has_wd = need_wd
        return self.wf, self.wd


def mark_params_updated(params):
    """Call after a kernel modified parameters through raw pointers (fused SGD/EMA)."""
    for p in params:
        p._ustrun_epoch = getattr(p, "_ustrun_epoch", 0) + 1


def _impl_for(cin, cout, dtype_code):
    if dtype_code == L.BF16 and not _FORCE_SIMT and cin % 64 == 0 and cout % 64 == 0:
        return L.TCGEN05
    return L.SIMT


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


# --------------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------------
def input_nchw(x: torch.Tensor) -> Act:
    """NCHW fp32 module input -> NHWC activation (custom_transforms.py:739-747 produces NCHW)."""
    if x.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
    L.require_device()
    x = _f32(x).contiguous()
    B, C, H, W = x.shape
    a = Act.new(B, H, W, C, device=x.device)
    _call("ustrun_nchw_to_nhwc", _ptr(x), a.ptr, a.dtype_code, B, C, H, W, a.ld, _stream())
    return a


def to_nchw(a: Act) -> torch.Tensor:
    out = torch.empty((a.B, a.C, a.H, a.W), dtype=torch.float32, device=a.t.device)
    _call("ustrun_nhwc_to_nchw", a.ptr, a.dtype_code, a.ld, _ptr(out), a.B, a.C, a.H, a.W, _stream())
    return out


def _raw_conv(x: Act, wpk, bias, y: Act, ks, partials=None, out_nchw=None):
    nparts = ctypes.c_int(0)
    impl = L.SIMT if out_nchw is not None else _impl_for(x.C, y.C if y is not None else 0, x.dtype_code)
    cout = out_nchw.shape[1] if out_nchw is not None else y.C
    _profiled("tc_conv" if impl == L.TCGEN05 else "simt_conv", 2.0 * x.npix * x.C * cout * ks * ks, "ustrun_conv_fwd", impl, x.ptr, x.ld, _ptr(wpk), _ptr(bias), _ptr(out_nchw) if out_nchw is not None else y.ptr,
          0 if out_nchw is not None else y.ld, x.dtype_code, x.B, x.H, x.W, x.C, cout, ks, 1 if out_nchw is not None else 0,
          _ptr(partials), ctypes.byref(nparts), _stream(), meta=(x.B, x.H, x.W, x.C, cout, ks))
    return nparts.value


def _wgrad(dy: Act, x: Act, dw: torch.Tensor, accumulate: int, ks: int):
    impl = _impl_for(x.C, dy.C, x.dtype_code)
    nbytes = L.lib.ustrun_conv_wgrad_workspace_bytes(impl, x.B, x.H, x.W, x.C, dy.C, ks)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=x.t.device)
    _profiled("tc_wgrad" if impl == L.TCGEN05 else "simt_wgrad", 2.0 * x.npix * x.C * dy.C * ks * ks,
              "ustrun_conv_wgrad", impl, dy.ptr, dy.ld, x.ptr, x.ld, _ptr(dw), accumulate, x.dtype_code, x.B, x.H, x.W, x.C, dy.C, ks,
          _ptr(ws), int(nbytes), _stream(), meta=(x.B, x.H, x.W, x.C, dy.C, ks))


class BNState:
    """The tensors of one nn.BatchNorm2d (weight, bias, running_mean, running_var, num_batches_tracked)."""

    def __init__(self, bn):
        self.bn = bn

    @property
    def eps(self):
        return float(self.bn.eps)

    @property
    def momentum(self):
        return 0.1 if self.bn.momentum is None else float(self.bn.momentum)


def conv_bn_act(ctx: Ctx, x: Act, conv, bn, act: int, out: Optional[Act] = None, pool: bool = False, packed: PackedConv = None):
    """conv(k=1|3, pad=k//2) -> BatchNorm2d (train: batch stats, eval: running stats) -> activation
    [-> MaxPool2d(2)].  Returns (y, pooled|None).  ``out`` lets the result land in a concat buffer.
    Reference: unet_parts.py:15-21,34 / unet.py:52-72,96-117."""
    ks = conv.kernel_size[0]
    cout = conv.out_channels
    wf, wd = packed.get(conv.weight, need_wd=ctx.need_grad or conv.weight.requires_grad)
    dev = x.t.device
    raw = x.like(cout)
    training = ctx.training and bn.training if hasattr(bn, "training") else ctx.training
    use_batch_stats = training or (bn.running_mean is None)
    partials = torch.empty(L.MAX_PARTS * 2 * cout, dtype=torch.float32, device=dev) if use_batch_stats else None
    nparts = _raw_conv(x, wf, None, raw, ks, partials)
    stats = torch.empty(4 * cout, dtype=torch.float32, device=dev)      # scale, shift, mean, rstd
    scale, shift, mean, rstd = stats[:cout], stats[cout:2 * cout], stats[2 * cout:3 * cout], stats[3 * cout:]
    count = float(x.npix)
    track = bn.track_running_stats and bn.running_mean is not None and training
    peer = ctx.bn_peer if use_batch_stats else None
    stat_ptr = ctx.stats.slot_ptr(bn, ctx.slot) if (track and ctx.stats is not None and (ctx.bn_sync is None or peer is not None)) else None
    if use_batch_stats and ctx.bn_sync is not None and peer is None:
        sums = torch.empty(2 * cout, dtype=torch.float32, device=dev)
        _call("ustrun_bn_reduce_partials", _ptr(partials), nparts, cout, _ptr(sums), _stream())
        ctx.bn_sync(sums)
        partials, nparts, count = sums, 1, count * ctx.bn_world
    if peer is not None:
        inpl = track and stat_ptr is None
        pargs = peer.args()                    # sequence number taken in PROGRAM order
        def fin_fn():
            _call("ustrun_bn_finalize_peer", _ptr(partials), nparts, cout, count * ctx.bn_world, _ptr(bn.weight), _ptr(bn.bias), _ptr(conv.bias),
                  _ptr(bn.running_mean) if inpl else None, _ptr(bn.running_var) if inpl else None,
                  _ptr(bn.num_batches_tracked) if inpl else None, 0.1 if bn.momentum is None else float(bn.momentum), float(bn.eps),
                  _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), stat_ptr, *pargs, _stream())
        if ctx.grads_on_side:
            on_sync_stream(fin_fn)
        else:
            fin_fn()
    else:
      inplace = (track and stat_ptr is None) or not use_batch_stats
      _call("ustrun_bn_finalize", _ptr(partials), nparts, cout, count, _ptr(bn.weight), _ptr(bn.bias), _ptr(conv.bias),
          _ptr(bn.running_mean) if inplace else None, _ptr(bn.running_var) if inplace else None,
          _ptr(bn.num_batches_tracked) if (track and stat_ptr is None) else None, 0.1 if bn.momentum is None else float(bn.momentum), float(bn.eps),
          1 if use_batch_stats else 0, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), stat_ptr, _stream())
    y = out if out is not None else x.like(cout)
    pooled = Act.new(x.B, x.H // 2, x.W // 2, cout, dtype=x.t.dtype, device=dev) if pool else None
    _profiled("hbm_bn_act", raw.npix * cout * raw.t.element_size() * (2.25 if pool else 2.0), "ustrun_bn_act_fwd", raw.ptr, raw.ld, _ptr(scale), _ptr(shift), act, y.ptr, y.ld, pooled.ptr if pool else None,
          pooled.ld if pool else 0, x.dtype_code, x.B, x.H, x.W, cout, _stream())

    if ctx.need_grad:
        if not use_batch_stats:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the SSL step")
        world = ctx.bn_world if ctx.bn_sync is not None else 1
        bn_sync = ctx.bn_sync
        bn_peer = ctx.bn_peer
        gos = ctx.grads_on_side and bn_sync is None
        multi_lane = ctx.grads_on_side

        def bwd(sink: GradSink):
            G = y.g
            if pool and pooled.g is not None:
                Gn = y.like() if y.parent is None else Act.new(y.B, y.H, y.W, y.C, dtype=y.t.dtype, device=dev)
                _profiled("hbm_maxpool_bwd", y.npix * y.C * y.t.element_size() * (3.25 if G is not None else 2.25), "ustrun_maxpool_bwd", y.ptr, y.ld, pooled.g.ptr, pooled.g.ld, G.ptr if G is not None else None,
                      G.ld if G is not None else 0, Gn.ptr, Gn.ld, y.dtype_code, y.B, y.H, y.W, y.C, _stream())
                G = Gn
            if G is None:
                return
            part = torch.empty(L.MAX_PARTS * 2 * cout, dtype=torch.float32, device=dev)
            np_ = ctypes.c_int(0)
            _profiled("hbm_bn_bwd_reduce", raw.npix * cout * raw.t.element_size() * 2.0, "ustrun_bn_bwd_reduce", G.ptr, G.ld, raw.ptr, raw.ld, _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), act,
                  raw.dtype_code, raw.npix, cout, _ptr(part), ctypes.byref(np_), _stream())
            n_parts, cnt = np_.value, float(raw.npix)
            coef = torch.empty(3 * cout, dtype=torch.float32, device=dev)
            dg, acc_g = sink.get(bn.weight) if bn.weight is not None else (None, 0)
            db, acc_b = sink.get(bn.bias) if bn.bias is not None else (None, 0)
            if bn_peer is not None:
                pargs = bn_peer.args()
                def bfin_fn(part=part, n_parts=n_parts, cnt=cnt):
                    _call("ustrun_bn_bwd_finalize_peer", _ptr(part), n_parts, cout, cnt * world, _ptr(bn.weight), _ptr(rstd), _ptr(dg), _ptr(db),
                          1 if (acc_g or acc_b) else 0, _ptr(coef), *pargs, _stream())
                    sink.done(bn.weight, bn.bias)
                if multi_lane:
                    on_sync_stream(bfin_fn)      # dgamma / dbeta of all lanes accumulate on this one stream, in program order
                else:
                    bfin_fn()
            elif bn_sync is not None:
                sums = torch.empty(2 * cout, dtype=torch.float32, device=dev)
                _call("ustrun_bn_reduce_partials", _ptr(part), n_parts, cout, _ptr(sums), _stream())
                bn_sync(sums)
                part, n_parts, cnt = sums, 1, cnt * world
            if bn_peer is None and gos:
                # coefficients on this lane's stream; dgamma / dbeta accumulate on the weight-gradient stream, whose program
                # order (branch after branch) keeps the accumulation order of the single-lane step
                _call("ustrun_bn_bwd_finalize", _ptr(part), n_parts, cout, cnt, _ptr(bn.weight), _ptr(rstd), None, None, 0, 1.0, _ptr(coef), _stream())
                def bn_grads_fn(part=part, n_parts=n_parts, cnt=cnt, dg=dg, db=db, acc=1 if (acc_g or acc_b) else 0):
                    _call("ustrun_bn_bwd_finalize", _ptr(part), n_parts, cout, cnt, _ptr(bn.weight), _ptr(rstd), _ptr(dg), _ptr(db), acc, 1.0, None, _stream())
                    sink.done(bn.weight, bn.bias)
                on_side_stream(bn_grads_fn, (part, rstd))
            elif bn_peer is None:
              _call("ustrun_bn_bwd_finalize", _ptr(part), n_parts, cout, cnt, _ptr(bn.weight), _ptr(rstd), _ptr(dg), _ptr(db),
                  1 if (acc_g or acc_b) else 0, 1.0 / world if bn_sync is not None else 1.0, _ptr(coef), _stream())
            if bn_peer is None and not gos:
                sink.done(bn.weight, bn.bias)
            draw = raw.like()
            _profiled("hbm_bn_bwd_apply", raw.npix * cout * raw.t.element_size() * 3.0, "ustrun_bn_bwd_apply", G.ptr, G.ld, raw.ptr, raw.ld, _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), _ptr(coef),
                  act, draw.ptr, draw.ld, raw.dtype_code, raw.npix, cout, _stream())
            if conv.bias is not None:
                dbias, acc = sink.get(conv.bias)          # BN removes the mean: d/dbias == 0 exactly
                if not acc:
                    dbias.zero_()
                sink.done(conv.bias)
            def wgrad_fn():
                dw, acc_w = sink.get(conv.weight)
                _wgrad(draw, x, dw, acc_w, ks)
                sink.done(conv.weight)
            if not _WGRAD_AFTER_DGRAD:
                on_side_stream(wgrad_fn, (draw.t, x.t))
            if x.needs_grad:
                gx = Act.new(x.B, x.H, x.W, x.C, dtype=x.t.dtype, device=dev)
                _raw_conv(draw, wd, None, gx, ks)
                _assign_grad(x, gx)
            if _WGRAD_AFTER_DGRAD:
                on_side_stream(wgrad_fn, (draw.t, x.t))

        ctx.tape.append(bwd)
        y.needs_grad = True
        if pooled is not None:
            pooled.needs_grad = True
    return y, pooled


def _assign_grad(x: Act, gx: Act):
    if x.parent is not None:
        raise RuntimeError("gradient of a view must come from its parent buffer")
    x.g = gx


def conv_transpose2x2(ctx: Ctx, x: Act, up, out: Act, packed: PackedConv):
    """nn.ConvTranspose2d(k2, s2) + bias, written straight into ``out`` (usually the second half
    of the decoder's concat buffer): unet_parts.py:53,57,62-67."""
    wf, wd = packed.get(up.weight, transposed=True, need_wd=ctx.need_grad or up.weight.requires_grad)
    cin, cout = up.in_channels, up.out_channels
    impl = _impl_for(cin, cout, x.dtype_code)
    tcls = "tc_convT" if impl == L.TCGEN05 else "simt_conv"
    tflops = 2.0 * x.npix * cin * cout * 4
    _profiled(tcls, tflops, "ustrun_convT2x2_fwd", impl, x.ptr, x.ld, _ptr(wf), _ptr(up.bias), out.ptr, out.ld, x.dtype_code, x.B, x.H, x.W, cin, cout, _stream(),
              meta=(x.B, x.H, x.W, cin, cout, "T"))
    if ctx.need_grad:
        dev = x.t.device
        gos = ctx.grads_on_side

        def bwd(sink: GradSink):
            G = out.g
            if G is None:
                return
            def wgrad_fn():
                dw, acc = sink.get(up.weight)
                nbytes = L.lib.ustrun_conv_wgrad_workspace_bytes(impl, x.B, x.H, x.W, cin, cout, 2)
                ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
                _profiled("tc_wgradT" if impl == L.TCGEN05 else "simt_wgrad", tflops, "ustrun_convT2x2_wgrad", impl, G.ptr, G.ld, x.ptr, x.ld, _ptr(dw), acc, x.dtype_code,
                          x.B, x.H, x.W, cin, cout, _ptr(ws), int(nbytes), _stream(), meta=(x.B, x.H, x.W, cin, cout, "T"))
                sink.done(up.weight)
            on_side_stream(wgrad_fn, (G.t, x.t))
            if up.bias is not None:
                def bias_fn():
                    db, accb = sink.get(up.bias)
                    wsb = torch.empty(L.MAX_PARTS * cout, dtype=torch.float32, device=dev)
                    _call("ustrun_channel_sum", G.ptr, G.ld, G.dtype_code, G.npix, cout, _ptr(db), accb, _ptr(wsb), _stream())
                    sink.done(up.bias)
                if gos:
                    on_side_stream(bias_fn, (G.t,))
                else:
                    bias_fn()
            if x.needs_grad:
                gx = Act.new(x.B, x.H, x.W, x.C, dtype=x.t.dtype, device=dev)
                _profiled(tcls, tflops, "ustrun_convT2x2_dgrad", impl, G.ptr, G.ld, _ptr(wd), gx.ptr, gx.ld, x.dtype_code, x.B, x.H, x.W, cin, cout, _stream(),
                          meta=(x.B, x.H, x.W, cin, cout, "Td"))
                _assign_grad(x, gx)

        ctx.tape.append(bwd)
        out.needs_grad = True
    return out


def upsample2x(ctx: Ctx, x: Act, align_corners: bool, out: Optional[Act] = None):
    """nn.Upsample(scale_factor=2, mode='bilinear'): unet.py:84,127 / unet_parts.py:50."""
    y = out if out is not None else Act.new(x.B, 2 * x.H, 2 * x.W, x.C, dtype=x.t.dtype, device=x.t.device)
    _call("ustrun_upsample2x_fwd", x.ptr, x.ld, y.ptr, y.ld, x.dtype_code, x.B, x.H, x.W, x.C, 1 if align_corners else 0, _stream())
    if ctx.need_grad:
        def bwd(sink: GradSink):
            G = y.g
            if G is None or not x.needs_grad:
                return
            gx = Act.new(x.B, x.H, x.W, x.C, dtype=x.t.dtype, device=x.t.device)
            _call("ustrun_upsample2x_bwd", G.ptr, G.ld, gx.ptr, gx.ld, x.dtype_code, x.B, x.H, x.W, x.C, 1 if align_corners else 0, _stream())
            _assign_grad(x, gx)

        ctx.tape.append(bwd)
        y.needs_grad = x.needs_grad
    return y


def head_conv(ctx: Ctx, x: Act, conv, packed: PackedConv):
    """Logits head (OutConv 1x1: unet_parts.py:74; out1/seg1 3x3: unet.py:182,312): fp32 NCHW out.
    Returns (logits, backward_fn(dlogits, sink))."""
    ks = conv.kernel_size[0]
    cout = conv.out_channels
    wf, wd = packed.get(conv.weight, need_wd=ctx.need_grad or conv.weight.requires_grad)
    logits = torch.empty((x.B, cout, x.H, x.W), dtype=torch.float32, device=x.t.device)
    _raw_conv(x, wf, conv.bias, None, ks, out_nchw=logits)
    gos = ctx.grads_on_side

    def bwd(dlogits: torch.Tensor, sink: GradSink):
        dev = x.t.device
        dlogits = _f32(dlogits).contiguous()
        cpad = cout                                        # narrow NHWC copy of dlogits
        G = Act.new(x.B, x.H, x.W, cpad, dtype=x.t.dtype, device=dev)
        _call("ustrun_nchw_to_nhwc", _ptr(dlogits), G.ptr, G.dtype_code, x.B, cout, x.H, x.W, G.ld, _stream())
        def wgrad_fn():
            dw, acc = sink.get(conv.weight)
            nbytes = L.lib.ustrun_conv_wgrad_workspace_bytes(L.SIMT, x.B, x.H, x.W, x.C, cout, ks)
            ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            _call("ustrun_conv_wgrad", L.SIMT, G.ptr, G.ld, x.ptr, x.ld, _ptr(dw), acc, x.dtype_code, x.B, x.H, x.W, x.C, cout, ks,
                  _ptr(ws), int(nbytes), _stream())
            sink.done(conv.weight)
        on_side_stream(wgrad_fn, (G.t, x.t))
        if conv.bias is not None:
            def bias_fn():
                db, accb = sink.get(conv.bias)
                wsb = torch.empty(L.MAX_PARTS * cout, dtype=torch.float32, device=dev)
                _call("ustrun_channel_sum", G.ptr, G.ld, G.dtype_code, G.npix, cout, _ptr(db), accb, _ptr(wsb), _stream())
                sink.done(conv.bias)
            if gos:
                on_side_stream(bias_fn, (G.t,))
            else:
                bias_fn()
        if x.needs_grad:
            gx = Act.new(x.B, x.H, x.W, x.C, dtype=x.t.dtype, device=dev)
            nparts = ctypes.c_int(0)
            _call("ustrun_conv_fwd", L.SIMT, G.ptr, G.ld, _ptr(wd), None, gx.ptr, gx.ld, x.dtype_code, x.B, x.H, x.W, cout, x.C, ks, 0,
                  None, ctypes.byref(nparts), _stream())
            _assign_grad(x, gx)

    return logits, bwd


def batchnorm_only(ctx: Ctx, x: Act, bn):
    """Stand-alone BatchNorm2d on an activation (used by DomainSpecificBatchNorm2d.forward when the
    module is called directly: dsbn.py:24-27).  Statistics come from the two-stage reduction kernel
    (bn_bwd_reduce with g=x, mean=0, rstd=1 computes sum x and sum x*x), then finalize + apply."""
    C, dev = x.C, x.t.device
    training = ctx.training
    use_batch = training or bn.running_mean is None
    stats = torch.empty(4 * C, dtype=torch.float32, device=dev)
    scale, shift, mean, rstd = stats[:C], stats[C:2 * C], stats[2 * C:3 * C], stats[3 * C:]
    part, nparts = None, 0
    if use_batch:
        ident = torch.empty(3 * C, dtype=torch.float32, device=dev)
        ident[:C] = 0.0
        ident[C:] = 1.0
        zero, one = ident[:C], ident[C:2 * C]
        part = torch.empty(L.MAX_PARTS * 2 * C, dtype=torch.float32, device=dev)
        np_ = ctypes.c_int(0)
        _call("ustrun_bn_bwd_reduce", x.ptr, x.ld, x.ptr, x.ld, _ptr(zero), _ptr(one), _ptr(one), _ptr(zero), L.ACT_NONE, x.dtype_code,
              x.npix, C, _ptr(part), ctypes.byref(np_), _stream())
        nparts = np_.value
    count = float(x.npix)
    if use_batch and ctx.bn_sync is not None:
        sums = torch.empty(2 * C, dtype=torch.float32, device=dev)
        _call("ustrun_bn_reduce_partials", _ptr(part), nparts, C, _ptr(sums), _stream())
        ctx.bn_sync(sums)
        part, nparts, count = sums, 1, count * ctx.bn_world
    track = bn.track_running_stats and bn.running_mean is not None and training
    _call("ustrun_bn_finalize", _ptr(part), nparts, C, count, _ptr(bn.weight), _ptr(bn.bias), None,
          _ptr(bn.running_mean) if (track or not use_batch) else None, _ptr(bn.running_var) if (track or not use_batch) else None,
          _ptr(bn.num_batches_tracked) if track else None, 0.1 if bn.momentum is None else float(bn.momentum), float(bn.eps),
          1 if use_batch else 0, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd), None, _stream())
    y = x.like()
    _call("ustrun_bn_act_fwd", x.ptr, x.ld, _ptr(scale), _ptr(shift), L.ACT_NONE, y.ptr, y.ld, None, 0, x.dtype_code, x.B, x.H, x.W, C, _stream())
    if ctx.need_grad:
        world, bn_sync = ctx.bn_world, ctx.bn_sync

        def bwd(sink: GradSink):
            G = y.g
            if G is None:
                return
            p2 = torch.empty(L.MAX_PARTS * 2 * C, dtype=torch.float32, device=dev)
            n2 = ctypes.c_int(0)
            _call("ustrun_bn_bwd_reduce", G.ptr, G.ld, x.ptr, x.ld, _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), L.ACT_NONE,
                  x.dtype_code, x.npix, C, _ptr(p2), ctypes.byref(n2), _stream())
            parts, n_parts, cnt = p2, n2.value, float(x.npix)
            if bn_sync is not None:
                sums2 = torch.empty(2 * C, dtype=torch.float32, device=dev)
                _call("ustrun_bn_reduce_partials", _ptr(p2), n_parts, C, _ptr(sums2), _stream())
                bn_sync(sums2)
                parts, n_parts, cnt = sums2, 1, cnt * world
            coef = torch.empty(3 * C, dtype=torch.float32, device=dev)
            dg, a1 = sink.get(bn.weight) if bn.weight is not None else (None, 0)
            db, a2 = sink.get(bn.bias) if bn.bias is not None else (None, 0)
            _call("ustrun_bn_bwd_finalize", _ptr(parts), n_parts, C, cnt, _ptr(bn.weight), _ptr(rstd), _ptr(dg), _ptr(db), 1 if (a1 or a2) else 0,
                  1.0 / world if bn_sync is not None else 1.0, _ptr(coef), _stream())
            sink.done(bn.weight, bn.bias)
            if x.needs_grad:
                gx = x.like()
                _call("ustrun_bn_bwd_apply", G.ptr, G.ld, x.ptr, x.ld, _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift), _ptr(coef),
                      L.ACT_NONE, gx.ptr, gx.ld, x.dtype_code, x.npix, C, _stream())
                _assign_grad(x, gx)

        ctx.tape.append(bwd)
        y.needs_grad = True
    return y
