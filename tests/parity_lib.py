"""Full-size parity harness (GPU, test infrastructure): the sm_100a SSL step against the oracle port executed by PyTorch
eager + cuDNN ON THE SAME B200 (TF32 off) in float64 (ground truth), fp32, bf16 autocast and fp16 autocast (the precision
the reference ships, train.py:30,551-552), at BASELINE.json's batch / image sizes and on WARMED weights.

Warm-up: the network is trained for N steps with OUR step (fast) on the learnable blob task of tests/synth_tasks.py; the
resulting student / teacher / momentum state is then just an input that every implementation steps from.  That is the
regime VERDICT r01 asked for: BatchNorm statistics of structured activations, a confident teacher (mask means 0.3-0.9)
instead of the degenerate random-init network.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

from synth_tasks import blob_batch, to_device_batch

PLANES = ("pseudo_label", "mask", "pseudo_label_w", "mask_w", "pseudo_label_ul", "mask_ul", "pseudo_label_lu", "mask_lu")
LOGIT_TAGS = ("t1", "t2", "t3", "s0", "lb", "ul", "lu", "s")


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten().to(a.device)
    return float((a - b).norm() / (b.norm() + 1e-300))


def make_pair(kind, c, k, seed=1337):
    """(student, teacher) drop-in modules on the GPU; kind: unet_a | unet_b [+ _dsbn3]."""
    kw = dict(norm="dsbn", num_domains=3) if kind.endswith("_dsbn3") else {}
    torch.manual_seed(seed)
    if kind.startswith("unet_a"):
        from networks.unet_model import UNet
    else:
        from networks.unet import UNet
    student, teacher = UNet(c, k, **kw), UNet(c, k, **kw)
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.detach_()
    return student.cuda().train(), teacher.cuda().train()


def oracle_forward(kind):
    from oracle import unet_ref as U
    dsbn = kind.endswith("_dsbn3")
    if kind.startswith("unet_a"):
        return (lambda s, x, dl: U.unet_a_forward(s, x, True, domain_label=dl)) if dsbn else (lambda s, x: U.unet_a_forward(s, x, True))
    return (lambda s, x, dl: U.unet_b_forward(s, x, True, domain_label=dl)) if dsbn else (lambda s, x: U.unet_b_forward(s, x, True))


def domains_for(kind, d_lb, d_ulb, Bl, Bu):
    if not kind.endswith("_dsbn3"):
        return None
    u, l = torch.full((Bu,), d_ulb, dtype=torch.long), torch.full((Bl,), d_lb, dtype=torch.long)
    return dict(t1=u, t2=u, t3=l, s0=u, lb=l, ul=u, lu=l, s=u, lq=u)


def export_state(module, dtype=None):
    sd = OrderedDict()
    for n, v in module.state_dict().items():
        v = v.detach().clone()
        sd[n] = v.to(dtype) if (dtype is not None and v.is_floating_point()) else v
    return sd


def export_momentum(trainer, dtype=None):
    names = [n for n, _ in trainer.model.named_parameters()]
    out = {}
    for i, (n, p) in enumerate(zip(names, trainer.params)):
        if trainer.opt.first[i]:
            out[n] = None
        else:
            b = trainer.opt.flat_buf[trainer.opt.offsets[i]: trainer.opt.offsets[i] + p.numel()].view(p.shape).detach().clone()
            out[n] = b.to(dtype) if dtype is not None else b
    return out


def grads_of(trainer):
    """name -> gradient of the last step (views of the flat buffer, cloned); None where the step produced none."""
    names = [n for n, _ in trainer.model.named_parameters()]
    return {n: (trainer.opt.grad_view(i).detach().clone() if trainer.opt.has_grad[i] else None) for i, n in enumerate(names)}


class Case:
    def __init__(self, name, kind, c, k, H, W, B, branch="softmax", d_lb=0, d_ulb=2, threshold=0.95, max_iterations=60000, iter0=30000):
        self.name, self.kind, self.c, self.k, self.H, self.W, self.B, self.branch = name, kind, c, k, H, W, B, branch
        self.d_lb, self.d_ulb, self.threshold, self.max_iterations, self.iter0 = d_lb, d_ulb, threshold, max_iterations, iter0

    def batch(self, seed):
        return blob_batch(self.c, self.k, self.H, self.W, self.B, self.B, seed=seed, branch=self.branch)

    def extra(self):
        return dict(domain_lb=self.d_lb, domain_ulb=self.d_ulb) if self.kind.endswith("_dsbn3") else {}

    def trainer(self, student, teacher, use_graph=False):
        from ustrun.step import SSLTrainer
        tr = SSLTrainer(student, teacher, n_classes=self.k, branch=self.branch, base_lr=0.03, max_iterations=self.max_iterations,
                        threshold=self.threshold, use_graph=use_graph)
        tr.iter_num = self.iter0
        tr.lr = 0.03 * (1.0 - (self.iter0 - 1) / self.max_iterations) ** 0.9 if self.iter0 > 0 else 0.03
        return tr


def warm(case, steps, precision="bf16", seed0=1000, log=None):
    """Train with the sm_100a step from the seeded init; returns (student, teacher, trainer).  For DSBN every third step
    swaps the domain pair so that all three domains' BatchNorms have seen data."""
    from ustrun import engine as E
    E.set_precision(precision)
    student, teacher = make_pair(case.kind, case.c, case.k)
    tr = case.trainer(student, teacher)
    pairs = [(case.d_lb, case.d_ulb), (1, case.d_lb), (case.d_ulb, 1)]
    for i in range(steps):
        b = to_device_batch(case.batch(seed0 + i))
        extra = case.extra()
        if extra and i < steps - 3:
            extra = dict(domain_lb=pairs[i % 3][0], domain_ulb=pairs[i % 3][1])
        out = tr.step({**b, **extra})
        if log is not None and (i % 10 == 0 or i == steps - 1):
            log(f"    warm {i}: loss {float(out['loss']):.4f} mask {float(out['mask'].float().mean()):.3f} mask_w {float(out['mask_w'].float().mean()):.3f}")
    torch.cuda.synchronize()
    return student, teacher, tr


def oracle_step(case, st_s, st_t, bufs, batch, it, lr, dtype=torch.float32, autocast=None, loss_scale=1.0, update=True):
    """One oracle step on the GPU from the given state (cloned and cast here).  Returns (out, student_after, teacher_after)."""
    from oracle import ssl_step_ref as S
    cast = lambda v: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()).cuda()
    s = OrderedDict((n, cast(v)) for n, v in st_s.items())
    t = OrderedDict((n, cast(v)) for n, v in st_t.items())
    bf = {n: (None if v is None else cast(v)) for n, v in bufs.items()}
    b = {n: (v.to(dtype) if v.is_floating_point() else v).cuda() for n, v in batch.items()}
    dom = domains_for(case.kind, case.d_lb, case.d_ulb, case.B, case.B)
    with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
        out = S.ssl_step(oracle_forward(case.kind), s, t, bf, b, n_classes=case.k, branch=case.branch, iter_num=it, max_iterations=case.max_iterations,
                         lr=lr, threshold=case.threshold, domains=dom, loss_scale=loss_scale, update=update)
    return out, s, t


def ours_step(case, st_s, st_t, bufs, batch, it, lr, precision):
    """One sm_100a step from the given state; returns (out, grads, student_after, teacher_after)."""
    from ustrun import engine as E
    E.set_precision(precision)
    try:
        student, teacher = make_pair(case.kind, case.c, case.k)
        student.load_state_dict({n: v.float() for n, v in st_s.items()})
        teacher.load_state_dict({n: v.float() for n, v in st_t.items()})
        tr = case.trainer(student, teacher)
        tr.iter_num, tr.lr = it, lr
        names = [n for n, _ in student.named_parameters()]
        for i, n in enumerate(names):
            if bufs.get(n) is not None:
                p = tr.params[i]
                tr.opt.flat_buf[tr.opt.offsets[i]: tr.opt.offsets[i] + p.numel()].view(p.shape).copy_(bufs[n].float())
                tr.opt.first[i] = False
        out = tr.step({**to_device_batch(batch), **case.extra()}, keep_logits=True)
        torch.cuda.synchronize()
        return out, grads_of(tr), export_state(student), export_state(teacher)
    finally:
        E.set_precision("bf16")


def compare_step(ref_out, ref_s_after, st_s_before, got_out, got_grads, got_s_after):
    """Error summary of one implementation's step (got_*) against the float64 step (ref_*)."""
    r = {}
    r["loss"] = abs(float(got_out["loss"]) - float(ref_out["loss"])) / abs(float(ref_out["loss"]))
    for n in ("sup_loss", "unsup_loss_ul", "unsup_loss_lu", "unsup_loss_s"):
        r[n] = abs(float(got_out[n]) - float(ref_out[n])) / max(abs(float(ref_out[n])), 1e-12)
    r["logits"] = {t: rel(got_out["logits"][t], ref_out["logits"][t]) for t in LOGIT_TAGS}
    r["logits_max"] = max(r["logits"].values())
    planes = {}
    for p in PLANES:
        a, b = got_out[p].reshape(-1).long(), ref_out[p].reshape(-1).long().to(got_out[p].device)
        planes[p] = float((a == b).float().mean())
    r["planes_min_agreement"] = min(planes.values())
    r["planes"] = planes
    keys = [n for n, g in ref_out["grads"].items() if g is not None]
    gmax = max(float(ref_out["grads"][n].norm()) for n in keys)
    live = [n for n in keys if float(ref_out["grads"][n].norm()) > 1e-6 * gmax]
    cat = lambda G: torch.cat([G[n].detach().double().flatten().cuda() for n in live])
    r["grads_all"] = rel(cat(got_grads), cat(ref_out["grads"]))
    per = sorted(((rel(got_grads[n], ref_out["grads"][n]), n) for n in live), reverse=True)
    r["grads_worst"] = [(float(e), n) for e, n in per[:3]]
    r["grads_median"] = float(np.median([e for e, _ in per]))
    # the parameter UPDATE of the step (SGD with momentum + weight decay): ||d_ours - d_ref|| / ||d_ref||
    num = den = 0.0
    for n in live:
        d_ref = ref_s_after[n].double() - st_s_before[n].double().cuda()
        d_got = got_s_after[n].double().cuda() - st_s_before[n].double().cuda()
        num += float((d_got - d_ref).pow(2).sum())
        den += float(d_ref.pow(2).sum())
    r["update"] = (num / max(den, 1e-300)) ** 0.5
    rs = [rel(got_s_after[n], ref_s_after[n]) for n in ref_s_after if n.endswith(("running_mean", "running_var"))]
    r["running_stats_max"] = max(rs) if rs else 0.0
    return r


def oracle_as_got(out, s_after):
    """Adapt an oracle run (fp32 / autocast) to the (out, grads, state) triple compare_step takes."""
    return out, out["grads"], s_after
