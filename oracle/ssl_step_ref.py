"""Oracle (test infrastructure): torch-CPU restatement of one UST-RUN SSL training step.

All host randomness of the reference loop (CutMix boxes, ``choice``, the FFT-mixed images
``move_transx``, the confidence bank) enters as *inputs*; see SURVEY.md App. A.

Reference lines followed:
  * pseudo labels        ``train.py:643-667``  (``train_mnms.py:587-603``)
  * student argmax       ``train.py:668-674``
  * blends / ensemble    ``train.py:677-697``
  * forwards order       ``train.py:668,699-702``
  * losses               ``train.py:816-838`` + ``utils/losses.py:194-268`` + ``utils/ramps.py:19-26``
  * optimiser + EMA + lr ``train.py:512,840-856,87-93``
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# scalars
# --------------------------------------------------------------------------------------------
def sigmoid_rampup(current, rampup_length):
    """utils/ramps.py:19-26."""
    if rampup_length == 0:
        return 1.0
    current = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def consistency_weight(iter_num, max_iterations, consistency=1.0, consistency_rampup=200.0):
    """train.py:82-84 called as at train.py:819-820."""
    return consistency * sigmoid_rampup(iter_num // (max_iterations / consistency_rampup), consistency_rampup)


def ema_alpha(global_step, ema_decay=0.99):
    """train.py:91."""
    return min(1 - 1 / (global_step + 1), ema_decay)


def poly_lr(base_lr, iter_num, max_iterations):
    """train.py:854 (evaluated with the pre-increment iter_num, applied to the NEXT step)."""
    return base_lr * (1.0 - iter_num / max_iterations) ** 0.9


# --------------------------------------------------------------------------------------------
# DiceLossWithMask  (utils/losses.py:194-268)
# --------------------------------------------------------------------------------------------
def dice_loss_with_mask(inputs, target, n_classes, mask=None, weight=None, softmax=False,
                        sigmoid=False, multi=False, smooth=1e-10):
    if sigmoid and softmax:
        raise AssertionError
    if sigmoid:
        inputs = inputs.sigmoid()
        target = target.squeeze(1)
    elif softmax:
        inputs = torch.softmax(inputs, dim=1)

    def ratio(score, tgt, m=None):
        tgt = tgt.float()
        if m is None:
            i, y, z = (score * tgt).sum(), (tgt * tgt).sum(), (score * score).sum()
        else:
            m = m.float()
            i, y, z = (score * tgt * m).sum(), (tgt * tgt * m).sum(), (score * score * m).sum()
        return 1 - (2 * i + smooth) / (z + y + smooth)

    if multi:                                                   # losses.py:244-249
        return ratio(inputs, target, mask)
    onehot = torch.cat([(target == c).float() for c in range(n_classes)], dim=1)   # :199-205
    assert inputs.size() == onehot.size(), "predict & target shape do not match"
    weight = [1] * n_classes if weight is None else weight
    loss = 0.0
    if mask is not None:
        # :207-213 -- class c keeps pixels where mask*c == c, i.e. class 0 is NEVER masked
        mhot = torch.cat([(mask * c == c * torch.ones_like(mask)).float() for c in range(n_classes)], dim=1)
        for c in range(n_classes):
            loss = loss + ratio(inputs[:, c], onehot[:, c], mhot[:, c]) * weight[c]
    else:
        for c in range(n_classes):
            loss = loss + ratio(inputs[:, c], onehot[:, c]) * weight[c]
    return loss / n_classes


def masked_term(logits, target, mask, n_classes, branch):
    """One ``(ce*mask).mean() + dice`` term (train.py:816-817,829-836)."""
    if branch == "softmax":
        ce = F.cross_entropy(logits, target, reduction="none")
        if mask is not None:
            ce = ce * mask.squeeze(1)
        return ce.mean() + dice_loss_with_mask(logits, target.unsqueeze(1), n_classes, mask=mask, softmax=True)
    ce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    if mask is not None:
        # train.py:829: ce [B,2,H,W] * mask.squeeze(1) ([B,2,H,W] stays as is in the fundus branch)
        ce = ce * mask.squeeze(1)
    return ce.mean() + dice_loss_with_mask(logits, target.unsqueeze(1), n_classes, mask=mask, sigmoid=True, multi=True)


# --------------------------------------------------------------------------------------------
# pseudo labels and compositions
# --------------------------------------------------------------------------------------------
def pseudo_label(logits, threshold, branch):
    """train.py:649-667 for one logits tensor -> (label, mask)."""
    if branch == "softmax":
        prob, label = torch.max(torch.softmax(logits, dim=1), dim=1)
        return label, (prob > threshold).unsqueeze(1).float()
    prob = logits.sigmoid()
    return prob.ge(0.5).float(), prob.ge(threshold).float() + prob.le(1 - threshold).float()


def compose(t1, t2, t3, box, cut_label, cut_mask, choice, threshold, branch):
    """train.py:649-697 given the three teacher logits.  ``box`` is [Bu,H,W] float {0,1}.

    Returns dict with pseudo_label, mask, pseudo_label_w, mask_w, pseudo_label_ul, mask_ul,
    pseudo_label_lu, mask_lu (dtypes as the reference: int64 labels / fp32 masks for softmax,
    fp32 labels for the fundus branch)."""
    img_box = box.unsqueeze(1)
    label_box = box.unsqueeze(1) if branch == "sigmoid" else box
    pl, mask = pseudo_label(t1, threshold, branch)
    pl_ul_t, m_ul_t = pseudo_label(t2, threshold, branch)
    pl_lu_t, m_lu_t = pseudo_label(t3, threshold, branch)
    mask_w = m_ul_t * (1 - img_box) + m_lu_t * img_box                                   # :677
    pl_w = (pl_ul_t * (1 - label_box) + pl_lu_t * label_box).long()                     # :679
    if branch == "sigmoid":
        pl_w = pl_w.float()
        ensemble = (pl_w == pl).float() * mask
    else:
        ensemble = (pl_w == pl).unsqueeze(1).float() * mask
    mask_w = mask_w.clone()
    mask_w[ensemble == 0] = 0                                                            # :685
    ratio_before = None
    cl, cm = cut_label[choice], cut_mask[choice]
    pl_ul = (pl * (1 - label_box) + cl * label_box).long()                               # :690
    pl_lu = (cl * (1 - label_box) + pl * label_box).long()                               # :693
    if branch == "sigmoid":
        pl_ul, pl_lu = pl_ul.float(), pl_lu.float()
    sel = img_box.expand(mask.shape)
    mask_ul = torch.where(sel == 1, cm, mask)                                            # :691
    mask_lu = torch.where(sel == 0, cm, mask)                                            # :697
    return dict(pseudo_label=pl, mask=mask, pseudo_label_w=pl_w, mask_w=mask_w,
                pseudo_label_ul=pl_ul, mask_ul=mask_ul, pseudo_label_lu=pl_lu, mask_lu=mask_lu)


def mix(a, b, img_box):
    """a outside the box, b inside (train.py:644,646,689,692)."""
    return a * (1 - img_box) + b * img_box


# --------------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------------
def sgd_step(params, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD(momentum=.9, weight_decay=1e-4) single group (train.py:512); params with
    ``grad is None`` are skipped (no weight decay either), buffers are created on first use."""
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        g = g.add(p, alpha=weight_decay)          # same op forms as torch/optim/sgd.py (single-tensor)
        if bufs.get(k) is None:
            bufs[k] = g.clone()
        else:
            bufs[k].mul_(momentum).add_(g, alpha=1)
        p.add_(bufs[k], alpha=-lr)


def ema_update(teacher_params, student_params, alpha):
    """train.py:92-93 (parameters only, never buffers)."""
    for k, t in teacher_params.items():
        t.mul_(alpha).add_(student_params[k], alpha=1 - alpha)


def ssl_step(forward, student, teacher, bufs, batch, *, n_classes, branch="softmax", iter_num=0,
             max_iterations=30000, lr=0.03, base_lr=0.03, threshold=0.95, consistency=1.0,
             consistency_rampup=200.0, ema_decay=0.99, momentum=0.9, weight_decay=1e-4,
             lq=None, update=True, domains=None, loss_scale=1.0):
    """One step.  ``forward(state, x) -> logits`` runs a train-mode forward with BN side effects on
    ``state`` (a flat dict holding params *and* buffers).  ``student``/``teacher`` are such dicts,
    ``bufs`` the SGD momentum buffers (dict name -> tensor|None).

    batch keys: lb_x, lb_mask, ulb_w, ulb_s, move_transx, box [Bu,H,W], choice (LongTensor [Bu]),
    cut_img, cut_label, cut_mask.  ``lq`` (optional) = the batch-1 low-quality image whose forward
    only updates the student's BN running statistics (train.py:740).
    ``domains`` (DSBN networks only, not upstream): dict forward tag -> ``domain_label`` tensor; ``forward`` is then
    called as ``forward(state, x, domain_label)``.  Tags in call order: t1 t2 t3 (teacher), s0, lb, ul, lu, s, lq.
    ``loss_scale``: static GradScaler factor (train.py:551,842-845): the scaled loss is back-propagated and the
    gradients are unscaled before SGD -- a no-op in exact arithmetic, needed for the fp16-autocast yardstick.
    Returns a dict of losses / compositions / grads; mutates student, teacher, bufs if ``update``."""
    from .unet_ref import split_state
    b = batch
    if domains is not None:
        fwd0 = forward
        order = iter(("t1", "t2", "t3", "s0", "lb", "ul", "lu", "s", "lq"))
        forward = lambda state, x: fwd0(state, x, domains[next(order)])
    img_box = b["box"].unsqueeze(1)
    mix_img = b["cut_img"][b["choice"]]
    with torch.no_grad():                                                                 # :638-647
        t1 = forward(teacher, b["ulb_w"])
        t2 = forward(teacher, mix(b["ulb_w"], mix_img, img_box))
        t3 = forward(teacher, mix(mix_img, b["ulb_w"], img_box))
        comp = compose(t1, t2, t3, b["box"], b["cut_label"], b["cut_mask"], b["choice"], threshold, branch)
    params, _ = split_state(student)
    for p in params.values():
        p.requires_grad_(True)
        p.grad = None
    s0 = forward(student, b["ulb_w"])                                                     # :668
    stu_pl = pseudo_label(s0.detach(), threshold, branch)[0]
    x_ul = mix(b["ulb_s"], b["move_transx"], img_box)                                     # :689
    x_lu = mix(b["move_transx"], b["ulb_s"], img_box)                                     # :692
    l_lb = forward(student, b["lb_x"])                                                    # :699-702
    l_ul = forward(student, x_ul)
    l_lu = forward(student, x_lu)
    l_s = forward(student, b["ulb_s"])
    if lq is not None:
        forward(student, lq)                                                              # :740
    sup = masked_term(l_lb, b["lb_mask"], None, n_classes, branch)
    cw = consistency_weight(iter_num, max_iterations, consistency, consistency_rampup)
    ul = masked_term(l_ul, comp["pseudo_label_ul"], comp["mask_ul"], n_classes, branch)
    lu = masked_term(l_lu, comp["pseudo_label_lu"], comp["mask_lu"], n_classes, branch)
    s = masked_term(l_s, comp["pseudo_label_w"], comp["mask_w"], n_classes, branch)
    loss = sup + cw * (ul + lu + cw * s)                                                  # :838
    (loss * loss_scale if loss_scale != 1.0 else loss).backward()
    inv = 1.0 / loss_scale
    grads = {k: ((p.grad.detach().clone() if loss_scale == 1.0 else p.grad.detach() * inv) if p.grad is not None else None) for k, p in params.items()}
    for p in params.values():
        p.requires_grad_(False)
    out = dict(comp)
    out.update(loss=loss.detach(), sup_loss=sup.detach(), unsup_loss_ul=ul.detach(),
               unsup_loss_lu=lu.detach(), unsup_loss_s=s.detach(), consistency_weight=cw,
               stu_pseudo_label=stu_pl, grads=grads,
               logits=dict(t1=t1, t2=t2, t3=t3, s0=s0.detach(), lb=l_lb.detach(), ul=l_ul.detach(),
                           lu=l_lu.detach(), s=l_s.detach()))
    if update:
        with torch.no_grad():
            sgd_step(params, grads, bufs, lr, momentum, weight_decay)
            tparams, _ = split_state(teacher)
            ema_update(tparams, params, ema_alpha(iter_num, ema_decay))
        out["next_lr"] = poly_lr(base_lr, iter_num, max_iterations)
    return out


# --------------------------------------------------------------------------------------------
# synthetic inputs; shared by tests and bench so both sides see identical tensors
# --------------------------------------------------------------------------------------------
def cutmix_box(img_h, img_w, rng, size_min=0.02, size_max=0.4, ratio_1=0.3, ratio_2=1 / 0.3):
    """train.py:222-240 with p=1 (default --cutmix_prob) and a numpy Generator as the RNG."""
    box = np.zeros((img_h, img_w), np.float32)
    size = rng.uniform(size_min, size_max) * img_h * img_w
    while True:
        ratio = rng.uniform(ratio_1, ratio_2)
        cw, ch = int(np.sqrt(size / ratio)), int(np.sqrt(size * ratio))
        x, y = int(rng.integers(0, img_w)), int(rng.integers(0, img_h))
        if x + cw <= img_w and y + ch <= img_h:
            break
    box[y:y + ch, x:x + cw] = 1
    return box


def synthetic_batch(n_channels, n_classes, H, W, B_l, B_u, seed=1337, branch="softmax", bank=0):
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    u = lambda *s: torch.rand(*s, generator=g) * 2 - 1
    lb_x, ulb_w, move = u(B_l, n_channels, H, W), u(B_u, n_channels, H, W), u(B_u, n_channels, H, W)
    ulb_s = (ulb_w + 0.1 * torch.randn(B_u, n_channels, H, W, generator=g)).clamp(-1, 1)
    if branch == "softmax":
        lb_mask = torch.randint(0, n_classes, (B_l, H, W), generator=g)
        cut_label = lb_mask.clone()
        cut_mask = torch.ones(B_l, 1, H, W)
        if bank:
            cut_label = torch.cat([cut_label, torch.randint(0, n_classes, (bank, H, W), generator=g)])
            cut_mask = torch.cat([cut_mask, (torch.rand(bank, 1, H, W, generator=g) > 0.4).float()])
    else:
        lb_mask = torch.randint(0, 2, (B_l, n_classes, H, W), generator=g).float()
        cut_label = lb_mask.clone()
        cut_mask = torch.ones(B_l, n_classes, H, W)
        if bank:
            cut_label = torch.cat([cut_label, torch.randint(0, 2, (bank, n_classes, H, W), generator=g).float()])
            cut_mask = torch.cat([cut_mask, (torch.rand(bank, n_classes, H, W, generator=g) > 0.4).float()])
    cut_img = lb_x.clone()
    if bank:
        cut_img = torch.cat([cut_img, u(bank, n_channels, H, W)])
    box = torch.from_numpy(np.stack([cutmix_box(H, W, rng) for _ in range(B_u)]))
    choice = torch.from_numpy(rng.integers(0, cut_img.shape[0], B_u)).long()
    return dict(lb_x=lb_x, lb_mask=lb_mask, ulb_w=ulb_w, ulb_s=ulb_s, move_transx=move, box=box,
                choice=choice, cut_img=cut_img, cut_label=cut_label, cut_mask=cut_mask)
