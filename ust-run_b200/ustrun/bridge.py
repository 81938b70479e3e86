"""torch.autograd bridge: runs an engine program (a function of ``(Ctx, Act)`` built from
``engine`` ops) as ONE autograd node whose inputs are the module input and every parameter.

This is what lets ``train.py`` / ``train_mnms.py`` keep calling ``model(x)``, ``loss.backward()``,
``optimizer.step()`` unchanged (Tier A): logits come back as fp32 NCHW tensors that
torch ops can consume, gradients arrive as ``.grad`` on the fp32 nn.Parameters.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import engine as E

# cross-rank BatchNorm statistics hook, installed by ustrun.dp (None = per-rank statistics)
BN_SYNC = None
BN_WORLD = 1


class Feature:
    """Marks a program output that is returned as a non-differentiable NCHW copy."""

    def __init__(self, act):
        self.act = act


class _Program(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, program, training, need_grad, precision, x, *params):
        ectx = E.Ctx(training, need_grad, bn_sync=BN_SYNC)
        ectx.bn_world = BN_WORLD
        ctx.precision = precision
        with E.precision_scope(precision):
            a = E.input_nchw(x)
            a.needs_grad = bool(need_grad and x.requires_grad)
            outs = program(ectx, a)
            return _Program._finish_forward(ctx, ectx, outs, params, a)

    @staticmethod
    def _finish_forward(ctx, ectx, outs, params, a):
        results, items, nondiff = [], [], []
        for o in outs:
            if isinstance(o, tuple):                      # (logits, head backward)
                results.append(o[0])
                items.append(("head", o[1]))
            elif isinstance(o, Feature):
                t = E.to_nchw(o.act)
                results.append(t)
                nondiff.append(t)
                items.append(("none", None))
            else:                                         # differentiable activation output
                results.append(E.to_nchw(o))
                items.append(("act", o))
        if nondiff:
            ctx.mark_non_differentiable(*nondiff)
        ctx.ectx, ctx.items, ctx.params, ctx.input_act = ectx, items, params, a
        return tuple(results)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, *grads):
        if ctx.ectx is None:
            raise RuntimeError("backward through a ustrun program twice (the saved activations were released)")
        sink = E.GradSink()
        with E.precision_scope(ctx.precision):
            for (kind, obj), g in zip(ctx.items, grads):
                if g is None or kind == "none":
                    continue
                if kind == "head":
                    obj(g, sink)
                else:
                    obj.g = E.input_nchw(g)
            ctx.ectx.backward(sink)
            gx = None
            if ctx.input_act.needs_grad and ctx.input_act.g is not None:
                gx = E.to_nchw(ctx.input_act.g)
        out = tuple(sink.fresh.get(id(p)) for p in ctx.params)
        ctx.ectx = None                                   # release saved activations
        return (None, None, None, None, gx) + out


def run_program(module, program, x, training=None):
    """Execute ``program`` for ``module`` on NCHW input ``x``; returns a tuple of tensors."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("ustrun modules run on CUDA tensors only (there is no CPU path); got a "
                           + ("CPU tensor" if isinstance(x, torch.Tensor) else type(x).__name__))
    params = tuple(module.parameters())
    training = module.training if training is None else training
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    return _Program.apply(program, training, need_grad, E.tier_a_precision(), x, *params)
