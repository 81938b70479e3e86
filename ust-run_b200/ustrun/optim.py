"""Fused multi-tensor SGD(momentum, weight decay) + EMA teacher update (K15+K16).

Replaces ``optim.SGD(...).step()`` (train.py:512,848) and ``update_ema_variables`` (train.py:87-93,
851) by ONE kernel launch over all parameter tensors.  Gradients live in one flat fp32 buffer
(``grad_view(i)`` are views into it), which is also what the data-parallel all-reduce operates on.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib as L
from .engine import _call, _profiled, _ptr, _stream, mark_params_updated


class FusedSGDEMA:
    def __init__(self, params, ema_params=None, momentum=0.9, weight_decay=1e-4):
        self.params = list(params)
        self.ema = list(ema_params) if ema_params is not None else [None] * len(self.params)
        assert len(self.ema) == len(self.params)
        L.require_device()
        dev = self.params[0].device
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        # flat gradient + momentum buffers, each tensor 16-byte aligned inside them
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.offsets, self.total = offs, total
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(total, dtype=torch.float32, device=dev)
        self.first = [True] * len(self.params)
        self.has_grad = [True] * len(self.params)
        self._grad_by_id = {id(p): self.grad_view(i) for i, p in enumerate(self.params)}
        blk_t, blk_o = [], []
        for i, p in enumerate(self.params):
            for o in range(0, p.numel(), L.OPT_CHUNK):
                blk_t.append(i)
                blk_o.append(o)
        self.nblocks = len(blk_t)
        self.blk_tensor = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self.blk_offset = torch.tensor(blk_o, dtype=torch.int64, device=dev)
        self._tables = {}

    # -- checkpointing: same layout as torch.optim.SGD.state_dict() (utils/util.py:259-273 saves it, train.py:544 restores it)
    def state_dict(self):
        state = {}
        for i, p in enumerate(self.params):
            if not self.first[i]:
                state[i] = {"momentum_buffer": self.flat_buf[self.offsets[i]: self.offsets[i] + p.numel()].view(p.shape).clone()}
        group = {"lr": None, "momentum": self.momentum, "dampening": 0, "weight_decay": self.weight_decay, "nesterov": False,
                 "maximize": False, "foreach": None, "differentiable": False, "fused": None, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        self.momentum = float(groups[0].get("momentum", self.momentum))
        self.weight_decay = float(groups[0].get("weight_decay", self.weight_decay))
        self.first = [True] * len(self.params)
        self.flat_buf.zero_()
        for pos, key in enumerate(order):
            st = sd["state"].get(key, sd["state"].get(str(key)))
            if st is not None and st.get("momentum_buffer") is not None:
                p = self.params[pos]
                self.flat_buf[self.offsets[pos]: self.offsets[pos] + p.numel()].view(p.shape).copy_(st["momentum_buffer"])
                self.first[pos] = False
        self._tables = {}

    def grad_view(self, i):
        p = self.params[i]
        return self.flat_grad[self.offsets[i]: self.offsets[i] + p.numel()].view(p.shape)

    def grad_for(self, param):
        return self._grad_by_id[id(param)]

    def set_has_grad(self, i, flag):
        self.has_grad[i] = bool(flag)

    def zero_grad(self):
        self.flat_grad.zero_()

    def _table(self):
        """Device descriptor table for the current (first, has_grad) flags.  One cached table per flag set (DSBN: the set of
        BatchNorm tensors that received gradients depends on the domains of the step), uploaded from a PINNED host copy with a
        stream-ordered copy: legal inside a CUDA-graph capture, where it becomes a 3 KB memcpy node of the graph."""
        key = (tuple(self.first), tuple(self.has_grad), tuple(p.data_ptr() for p in self.params))
        ent = self._tables.get(key)
        if ent is None:
            dt = np.dtype([("p", "<u8"), ("g", "<u8"), ("buf", "<u8"), ("ema", "<u8"), ("n", "<i8"), ("first", "<i4"), ("pad", "<i4")])
            tab = np.zeros(len(self.params), dtype=dt)
            for i, p in enumerate(self.params):
                assert p.dtype == torch.float32 and p.is_contiguous()
                tab[i]["p"] = p.data_ptr()
                tab[i]["g"] = self.flat_grad.data_ptr() + 4 * self.offsets[i] if self.has_grad[i] else 0
                tab[i]["buf"] = self.flat_buf.data_ptr() + 4 * self.offsets[i]
                tab[i]["ema"] = self.ema[i].data_ptr() if self.ema[i] is not None else 0
                tab[i]["n"] = p.numel()
                tab[i]["first"] = 1 if self.first[i] else 0
            host = torch.from_numpy(tab.view(np.uint8).reshape(-1).copy()).pin_memory()
            dev = torch.empty(host.numel(), dtype=torch.uint8, device=self.flat_grad.device)
            dev.copy_(host, non_blocking=True)
            if len(self._tables) > 64:
                self._tables.clear()
            ent = self._tables[key] = (dev, host)
        return ent[0]

    def step(self, lr=None, alpha=None, grad_scale=1.0, do_sgd=True, hyper=None, do_ema=None):
        """SGD step with the gradients currently in the flat buffer, then (alpha given) EMA.
        ``hyper``: device tensor float32[>=3] = {lr, alpha, grad_scale}; the kernel then reads the per-step scalars from
        device memory (CUDA-graph replay) and ``lr`` / ``alpha`` / ``grad_scale`` are ignored (``do_ema`` selects the EMA half)."""
        table = self._table()
        nbytes = 28.0 * sum(p.numel() for p in self.params)
        if hyper is not None:
            _profiled("hbm_sgd_ema", nbytes, "ustrun_sgd_ema_multi_dev", _ptr(table), _ptr(self.blk_tensor), _ptr(self.blk_offset), self.nblocks, _ptr(hyper),
                      self.momentum, self.weight_decay, 1 if do_sgd else 0, 1 if (do_ema if do_ema is not None else True) else 0, _stream())
        else:
            _profiled("hbm_sgd_ema", nbytes, "ustrun_sgd_ema_multi", _ptr(table), _ptr(self.blk_tensor), _ptr(self.blk_offset), self.nblocks, float(lr), self.momentum,
                      self.weight_decay, float(alpha if alpha is not None else 0.0), float(grad_scale), 1 if do_sgd else 0,
                      1 if alpha is not None else 0, _stream())
        if do_sgd:
            for i in range(len(self.params)):
                if self.has_grad[i]:
                    self.first[i] = False
        mark_params_updated(self.params)
        mark_params_updated([e for e in self.ema if e is not None])
