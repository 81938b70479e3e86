"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference has no tests or golden vectors, so these fixtures -- outputs of
``/root/reference``'s own modules on seeded inputs -- are what pins the oracle.  The script also
asserts, while it runs, that the oracle restatement reproduces the reference bit-for-bit on CPU
(same torch ops in the same order), so a drift in either is caught at generation time.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("USTRUN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import ssl_step_ref as S  # noqa: E402
from oracle import unet_ref as U  # noqa: E402

from networks import unet as ref_unet_b  # noqa: E402  (reference)
from networks import unet_model as ref_unet_a  # noqa: E402
from networks.dsbn import DomainSpecificBatchNorm2d  # noqa: E402
from utils import losses as ref_losses  # noqa: E402
from utils import ramps as ref_ramps  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 1337


def np_(t):
    return t.detach().cpu().numpy()


def digest(state):
    """Small per-tensor fingerprint: (sum, abs-sum, first 4 values)."""
    out = {}
    for k, v in state.items():
        v = v.detach().double().flatten()
        out[k] = np.concatenate([[v.sum().item(), v.abs().sum().item()], np_(v[:4])]).astype(np.float64)
    return out


def assert_same_state(ref_mod, st):
    sd = ref_mod.state_dict()
    assert list(sd.keys()) == list(st.keys()), "state_dict keys/order differ"
    for k in sd:
        assert torch.equal(sd[k], st[k]), k


def run_model_case(name, ref_mod, st, fwd, x, n_classes, extra=None):
    """Train-mode forward + masked loss backward on the reference module and on the oracle."""
    g = torch.Generator().manual_seed(SEED + 7)
    tgt = torch.randint(0, n_classes, (x.shape[0],) + x.shape[2:], generator=g)
    msk = (torch.rand(x.shape[0], 1, *x.shape[2:], generator=g) > 0.3).float()
    ref_mod.train()
    ref_logits = ref_mod(x) if extra is None else ref_mod(x, **extra)
    dice = ref_losses.DiceLossWithMask(n_classes)
    ce = torch.nn.CrossEntropyLoss(reduction="none")
    ref_loss = (ce(ref_logits, tgt) * msk.squeeze(1)).mean() + dice(ref_logits, tgt.unsqueeze(1), mask=msk, softmax=True)
    ref_loss.backward()
    ref_grads = {k: p.grad for k, p in ref_mod.named_parameters()}

    params, _ = U.split_state(st)
    for p in params.values():
        p.requires_grad_(True)
    logits = fwd(st, x)
    loss = S.masked_term(logits, tgt, msk, n_classes, "softmax")
    loss.backward()
    assert torch.equal(logits, ref_logits), f"{name}: oracle logits != reference"
    assert torch.equal(loss, ref_loss), f"{name}: oracle loss != reference"
    for k, p in params.items():
        if ref_grads[k] is None:
            assert p.grad is None, k
        else:
            assert torch.allclose(p.grad, ref_grads[k], rtol=0, atol=0), f"{name}: grad {k}"
    assert_same_state(ref_mod, {k: v.detach() for k, v in st.items()})
    fx = {"x": np_(x), "target": np_(tgt), "mask": np_(msk), "logits": np_(ref_logits),
          "loss": np_(ref_loss)}
    for k, v in digest({k: v for k, v in ref_mod.state_dict().items()}).items():
        fx["state_after/" + k] = v
    for k, v in ref_grads.items():
        if v is not None:
            fx["grad/" + k] = np.concatenate([[v.double().norm().item(), v.double().sum().item()], np_(v.flatten()[:6]).astype(np.float64)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print(f"{name}: loss={ref_loss.item():.6f}  ok")


def case_unet_a(n_channels, n_classes, hw, batch):
    torch.manual_seed(SEED)
    ref = ref_unet_a.UNet(n_channels=n_channels, n_classes=n_classes)
    st = U.init_unet_a(n_channels, n_classes, seed=SEED)
    assert_same_state(ref, st)
    x = torch.rand(batch, n_channels, hw, hw, generator=torch.Generator().manual_seed(SEED + 1)) * 2 - 1
    run_model_case(f"unet_a_c{n_channels}_k{n_classes}_{hw}", ref, st, lambda s, t: U.unet_a_forward(s, t, True), x, n_classes)


def case_unet_b(n_channels, n_classes, hw, batch):
    torch.manual_seed(SEED)
    ref = ref_unet_b.UNet(n_channels=n_channels, n_classes=n_classes)
    st = U.init_unet_b(n_channels, n_classes, seed=SEED)
    assert_same_state(ref, st)
    x = torch.rand(batch, n_channels, hw, hw, generator=torch.Generator().manual_seed(SEED + 2)) * 2 - 1
    run_model_case(f"unet_b_c{n_channels}_k{n_classes}_{hw}", ref, st, lambda s, t: U.unet_b_forward(s, t, True), x, n_classes)


class _EncRec(torch.nn.Module):
    """The only DSBN network constructible upstream: Encoder(bn) + Rec_Decoder(dsbn)."""

    def __init__(self, c, classes, domains):
        super().__init__()
        self.enc = ref_unet_b.Encoder(c=c, norm="bn")
        self.dec = ref_unet_b.Rec_Decoder(num_classes=classes, norm="dsbn", num_domains=domains)

    def forward(self, x, domain_label=None):
        return self.dec(self.enc(x)[-1], domain_label=domain_label)


def case_dsbn(hw=32, batch=4, domains=3, classes=2, c=3):
    torch.manual_seed(SEED)
    ref = _EncRec(c, classes, domains)
    torch.manual_seed(SEED)
    enc = U.init_unet_b(c, classes, norm="bn", decoder=False)
    dec = U.init_rec_decoder(num_classes=classes, norm="dsbn", num_domains=domains)
    st = {**{"enc." + k: v for k, v in enc.items()}, **{"dec." + k: v for k, v in dec.items()}}
    assert_same_state(ref, st)
    x = torch.rand(batch, c, hw, hw, generator=torch.Generator().manual_seed(SEED + 3)) * 2 - 1
    dl = torch.tensor([2, 2, 0, 1])

    def fwd(s, t):
        e = {k[4:]: v for k, v in s.items() if k.startswith("enc.")}
        d = {k[4:]: v for k, v in s.items() if k.startswith("dec.")}
        return U.rec_decoder_forward(d, U.unet_b_encoder(e, t, True)[-1], dl, True)

    run_model_case("dsbn_encrec", ref, st, fwd, x, classes, extra={"domain_label": dl})
    # plain DSBN module semantics (dsbn.py:24-27): tuple return, bns[domain_label[0]] only
    torch.manual_seed(SEED)
    m = DomainSpecificBatchNorm2d(8, 3)
    y, dl2 = m(torch.ones(2, 8, 4, 4), torch.tensor([1, 0]))
    assert [int(b.num_batches_tracked) for b in m.bns] == [0, 1, 0] and dl2.tolist() == [1, 0]


def case_losses():
    fx = {}
    g = torch.Generator().manual_seed(SEED + 11)
    for C in (2, 3, 4):
        dice = ref_losses.DiceLossWithMask(C)
        logits = (torch.randn(2, C, 12, 10, generator=g) * 2).requires_grad_(True)
        tgt = torch.randint(0, C, (2, 12, 10), generator=g)
        msk = (torch.rand(2, 1, 12, 10, generator=g) > 0.5).float()
        for tag, m in (("m", msk), ("n", None)):
            ref = dice(logits, tgt.unsqueeze(1), mask=m, softmax=True)
            (gr,) = torch.autograd.grad(ref, logits)
            mine = S.dice_loss_with_mask(logits, tgt.unsqueeze(1), C, mask=m, softmax=True)
            assert torch.equal(ref, mine)
            fx[f"softmax_C{C}_{tag}/loss"], fx[f"softmax_C{C}_{tag}/grad"] = np_(ref), np_(gr)
        fx[f"softmax_C{C}/logits"], fx[f"softmax_C{C}/target"], fx[f"softmax_C{C}/mask"] = np_(logits), np_(tgt), np_(msk)
        # F8: class-0 dice term ignores the mask entirely
        allzero = dice(logits, tgt.unsqueeze(1), mask=torch.zeros_like(msk), softmax=True)
        fx[f"softmax_C{C}_zero_mask/loss"] = np_(allzero)
    dice = ref_losses.DiceLossWithMask(2)
    logits = (torch.randn(2, 2, 12, 10, generator=g) * 2).requires_grad_(True)
    tgt = torch.randint(0, 2, (2, 2, 12, 10), generator=g).float()
    msk = (torch.rand(2, 2, 12, 10, generator=g) > 0.5).float()
    for tag, m in (("m", msk), ("n", None)):
        ref = dice(logits, tgt.unsqueeze(1), mask=m, sigmoid=True, multi=True)
        (gr,) = torch.autograd.grad(ref, logits)
        assert torch.equal(ref, S.dice_loss_with_mask(logits, tgt.unsqueeze(1), 2, mask=m, sigmoid=True, multi=True))
        fx[f"sigmoid_{tag}/loss"], fx[f"sigmoid_{tag}/grad"] = np_(ref), np_(gr)
    fx["sigmoid/logits"], fx["sigmoid/target"], fx["sigmoid/mask"] = np_(logits), np_(tgt), np_(msk)
    for cur in (0, 1, 37.0, 199, 200, 500):
        assert S.sigmoid_rampup(cur, 200.0) == ref_ramps.sigmoid_rampup(cur, 200.0)
        fx[f"rampup/{cur}"] = np.float64(ref_ramps.sigmoid_rampup(cur, 200.0))
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **fx)
    print("losses: ok")


def case_step(model, branch, n_channels, n_classes, hw, B, iter_num, bank):
    """Whole step: the oracle glue driving REFERENCE modules/loss/SGD vs. the fully restated oracle."""
    torch.manual_seed(SEED)
    if model == "a":
        ref_s, ref_t = ref_unet_a.UNet(n_channels, n_classes), ref_unet_a.UNet(n_channels, n_classes)
        torch.manual_seed(SEED)
        st_s, st_t = U.init_unet_a(n_channels, n_classes), U.init_unet_a(n_channels, n_classes)
        fwd = lambda s, x: U.unet_a_forward(s, x, True)
    else:
        ref_s, ref_t = ref_unet_b.UNet(n_channels, n_classes), ref_unet_b.UNet(n_channels, n_classes)
        torch.manual_seed(SEED)
        st_s, st_t = U.init_unet_b(n_channels, n_classes), U.init_unet_b(n_channels, n_classes)
        fwd = lambda s, x: U.unet_b_forward(s, x, True)
    for p in ref_t.parameters():
        p.detach_()
    assert_same_state(ref_s, st_s), assert_same_state(ref_t, st_t)
    batch = S.synthetic_batch(n_channels, n_classes, hw, hw, B, B, seed=SEED, branch=branch, bank=bank)
    thr = 0.6 if branch == "softmax" else 0.55       # random-init nets are unconfident
    kw = dict(n_classes=n_classes, branch=branch, iter_num=iter_num, max_iterations=30000, lr=0.03,
              threshold=thr)
    # --- reference side: real modules + real DiceLossWithMask + torch.optim.SGD -------------
    ref_s.train(), ref_t.train()
    opt = torch.optim.SGD(ref_s.parameters(), lr=0.03, momentum=0.9, weight_decay=0.0001)
    dice = ref_losses.DiceLossWithMask(n_classes)
    ce = torch.nn.CrossEntropyLoss(reduction="none") if branch == "softmax" else torch.nn.BCEWithLogitsLoss(reduction="none")
    sm, sg, mu = (True, False, False) if branch == "softmax" else (False, True, True)
    b = batch
    img_box = b["box"].unsqueeze(1)
    mix_img = b["cut_img"][b["choice"]]
    with torch.no_grad():
        t1 = ref_t(b["ulb_w"]); t2 = ref_t(S.mix(b["ulb_w"], mix_img, img_box)); t3 = ref_t(S.mix(mix_img, b["ulb_w"], img_box))
        comp = S.compose(t1, t2, t3, b["box"], b["cut_label"], b["cut_mask"], b["choice"], thr, branch)
    # the reference composes mask_ul/mask_lu with boolean-mask assignment (train.py:688-697);
    # the oracle uses torch.where -- check the two formulations agree on this batch
    m_ul, m_lu = comp["mask"].clone(), comp["mask"].clone()
    sel = img_box.expand(m_ul.shape)
    m_ul[sel == 1] = b["cut_mask"][b["choice"]][sel == 1]
    m_lu[sel == 0] = b["cut_mask"][b["choice"]][sel == 0]
    assert torch.equal(m_ul, comp["mask_ul"]) and torch.equal(m_lu, comp["mask_lu"])
    ref_s(b["ulb_w"])
    outs = [ref_s(b["lb_x"]), ref_s(S.mix(b["ulb_s"], b["move_transx"], img_box)),
            ref_s(S.mix(b["move_transx"], b["ulb_s"], img_box)), ref_s(b["ulb_s"])]
    tg = [(b["lb_mask"], None), (comp["pseudo_label_ul"], comp["mask_ul"]),
          (comp["pseudo_label_lu"], comp["mask_lu"]), (comp["pseudo_label_w"], comp["mask_w"])]
    terms = []
    for o, (t, m) in zip(outs, tg):
        c = ce(o, t)
        if m is not None:
            c = c * m.squeeze(1)
        terms.append(c.mean() + dice(o, t.unsqueeze(1), mask=m, softmax=sm, sigmoid=sg, multi=mu))
    cw = 1.0 * ref_ramps.sigmoid_rampup(iter_num // (30000 / 200.0), 200.0)
    loss = terms[0] + cw * (terms[1] + terms[2] + cw * terms[3])
    opt.zero_grad(); loss.backward(); opt.step()
    alpha = min(1 - 1 / (iter_num + 1), 0.99)
    for ep, p in zip(ref_t.parameters(), ref_s.parameters()):
        ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)
    # --- oracle side ------------------------------------------------------------------------
    bufs = {}
    out = S.ssl_step(fwd, st_s, st_t, bufs, batch, **kw)
    assert torch.equal(out["loss"], loss.detach()), (out["loss"], loss)
    for k in comp:
        assert torch.equal(out[k], comp[k]), k
    assert_same_state(ref_s, {k: v.detach() for k, v in st_s.items()})
    assert_same_state(ref_t, {k: v.detach() for k, v in st_t.items()})
    name = f"step_{model}_{branch}_c{n_channels}_k{n_classes}_{hw}_b{B}_it{iter_num}_bank{bank}"
    fx = {"loss": np_(loss), "terms": np.array([t.item() for t in terms]), "cw": np.float64(cw),
          "threshold": np.float64(thr), "mask_mean": np.float64(comp["mask"].mean().item())}
    for k, v in comp.items():
        fx["comp/" + k] = np_(v).astype(np.uint8) if v.dtype in (torch.int64,) or v.max() <= 1 else np_(v)
    for k, v in out["logits"].items():
        fx["logits/" + k] = np_(v)
    for k, v in digest(ref_s.state_dict()).items():
        fx["student_after/" + k] = v
    for k, v in digest(ref_t.state_dict()).items():
        fx["teacher_after/" + k] = v
    for k, v in out["grads"].items():
        fx["grad/" + k] = np.concatenate([[v.double().norm().item(), v.double().sum().item()], np_(v.flatten()[:6]).astype(np.float64)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print(f"{name}: loss={loss.item():.6f} mask_mean={comp['mask'].mean().item():.3f} ok")



# ---- DSBN through the whole UNet-B: the reference modules under a documented run-time patch ----
# Upstream, ``UNet(norm='dsbn')`` raises TypeError at construction: ConvD/ConvU call ``normalization(planes, norm)``
# without ``num_domains`` (unet.py:38,82) and their forwards call ``self.bnX(x)`` without ``domain_label``
# (unet.py:58-70,98-115).  The patch below changes NOTHING in the reference's files; at run time it
#   (1) wraps ``networks.unet.normalization`` so that the 'dsbn' case receives ``num_domains`` (2 lines), and
#   (2) gives ``_DomainSpecificBatchNorm.forward`` a default ``domain_label`` taken from a context variable and
#       makes it return the bare tensor in that case (4 lines) -- the selection rule itself,
#       ``self.bns[domain_label[0]](x)`` (dsbn.py:26), is the reference's own line.
# ConvD / ConvU / UNet.forward, the ModuleList of BatchNorm2d and the initialisation then run unmodified.
class _DsbnPatch:
    current = [None]

    def __init__(self, num_domains):
        self.num_domains = num_domains

    def __enter__(self):
        from networks import dsbn as ref_dsbn
        self._norm, self._fwd = ref_unet_b.normalization, ref_dsbn._DomainSpecificBatchNorm.forward
        nd, orig = self.num_domains, self._norm
        ref_unet_b.normalization = lambda planes, norm='gn', num_domains=None: orig(planes, norm, nd if norm == 'dsbn' else num_domains)
        orig_fwd, cur = self._fwd, self.current

        def forward(mod, x, domain_label=None):
            if domain_label is not None:
                return orig_fwd(mod, x, domain_label)
            return orig_fwd(mod, x, cur[0])[0]
        ref_dsbn._DomainSpecificBatchNorm.forward = forward
        return self

    def __exit__(self, *exc):
        from networks import dsbn as ref_dsbn
        ref_unet_b.normalization, ref_dsbn._DomainSpecificBatchNorm.forward = self._norm, self._fwd


class _WithDomain(torch.nn.Module):
    """Calls the patched reference network with the context variable set (keeps run_model_case generic)."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x, domain_label=None):
        _DsbnPatch.current[0] = domain_label
        return self.net(x)

    def state_dict(self, *a, **k):
        return self.net.state_dict(*a, **k)

    def named_parameters(self, *a, **k):
        return self.net.named_parameters(*a, **k)


def case_unet_b_dsbn(n_channels=3, n_classes=2, hw=32, batch=2, domains=3):
    with _DsbnPatch(domains):
        torch.manual_seed(SEED)
        ref = ref_unet_b.UNet(n_channels=n_channels, n_classes=n_classes, norm="dsbn")
        st = U.init_unet_b(n_channels, n_classes, seed=SEED, norm="dsbn", num_domains=domains)
        assert_same_state(ref, st)
        x = torch.rand(batch, n_channels, hw, hw, generator=torch.Generator().manual_seed(SEED + 4)) * 2 - 1
        dl = torch.tensor([1, 2][:batch])
        run_model_case(f"unet_b_dsbn{domains}_c{n_channels}_k{n_classes}_{hw}", _WithDomain(ref), st,
                       lambda s, t: U.unet_b_forward(s, t, True, domain_label=dl), x, n_classes, extra={"domain_label": dl})


STEP_DOMAIN_TAGS = ("t1", "t2", "t3", "s0", "lb", "ul", "lu", "s")


def step_domains(d_lb, d_ulb, B):
    """Domain of each forward of the step (this repo's convention; upstream has no DSBN step): a forward belongs to
    the domain of the image that fills the area OUTSIDE the CutMix box -- t1/t2/s0/ul/s: the unlabelled batch's
    domain, t3/lb/lu: the labelled batch's."""
    u, l = torch.full((B,), d_ulb, dtype=torch.long), torch.full((B,), d_lb, dtype=torch.long)
    return dict(t1=u, t2=u, t3=l, s0=u, lb=l, ul=u, lu=l, s=u, lq=u)


def case_step_dsbn(n_channels=3, n_classes=2, hw=32, B=2, iter_num=3000, domains=3, d_lb=0, d_ulb=2):
    """Whole DSBN step: patched reference UNet-B modules + reference loss + torch SGD vs the restated oracle."""
    branch, thr = "softmax", 0.6
    with _DsbnPatch(domains):
        torch.manual_seed(SEED)
        ref_s = ref_unet_b.UNet(n_channels, n_classes, norm="dsbn")
        ref_t = ref_unet_b.UNet(n_channels, n_classes, norm="dsbn")
        torch.manual_seed(SEED)
        st_s = U.init_unet_b(n_channels, n_classes, norm="dsbn", num_domains=domains)
        st_t = U.init_unet_b(n_channels, n_classes, norm="dsbn", num_domains=domains)
        for p in ref_t.parameters():
            p.detach_()
        assert_same_state(ref_s, st_s), assert_same_state(ref_t, st_t)
        batch = S.synthetic_batch(n_channels, n_classes, hw, hw, B, B, seed=SEED, branch=branch)
        dom = step_domains(d_lb, d_ulb, B)
        ref_s.train(), ref_t.train()
        opt = torch.optim.SGD(ref_s.parameters(), lr=0.03, momentum=0.9, weight_decay=0.0001)
        dice = ref_losses.DiceLossWithMask(n_classes)
        ce = torch.nn.CrossEntropyLoss(reduction="none")
        b = batch
        img_box = b["box"].unsqueeze(1)
        mix_img = b["cut_img"][b["choice"]]

        def run(net, x, tag):
            _DsbnPatch.current[0] = dom[tag]
            return net(x)
        with torch.no_grad():
            t1 = run(ref_t, b["ulb_w"], "t1"); t2 = run(ref_t, S.mix(b["ulb_w"], mix_img, img_box), "t2"); t3 = run(ref_t, S.mix(mix_img, b["ulb_w"], img_box), "t3")
            comp = S.compose(t1, t2, t3, b["box"], b["cut_label"], b["cut_mask"], b["choice"], thr, branch)
        run(ref_s, b["ulb_w"], "s0")
        outs = [run(ref_s, b["lb_x"], "lb"), run(ref_s, S.mix(b["ulb_s"], b["move_transx"], img_box), "ul"),
                run(ref_s, S.mix(b["move_transx"], b["ulb_s"], img_box), "lu"), run(ref_s, b["ulb_s"], "s")]
        tg = [(b["lb_mask"], None), (comp["pseudo_label_ul"], comp["mask_ul"]), (comp["pseudo_label_lu"], comp["mask_lu"]), (comp["pseudo_label_w"], comp["mask_w"])]
        terms = []
        for o, (t, m) in zip(outs, tg):
            c = ce(o, t)
            if m is not None:
                c = c * m.squeeze(1)
            terms.append(c.mean() + dice(o, t.unsqueeze(1), mask=m, softmax=True))
        cw = 1.0 * ref_ramps.sigmoid_rampup(iter_num // (30000 / 200.0), 200.0)
        loss = terms[0] + cw * (terms[1] + terms[2] + cw * terms[3])
        opt.zero_grad(); loss.backward(); opt.step()
        alpha = min(1 - 1 / (iter_num + 1), 0.99)
        for ep, p in zip(ref_t.parameters(), ref_s.parameters()):
            ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)
        # only the BatchNorms of the two domains in play received gradients (SGD skips grad=None: no weight decay either)
        for n_, p in ref_s.named_parameters():
            if ".bns." in n_:
                d = int(n_.split(".bns.")[1].split(".")[0])
                assert (p.grad is not None) == (d in (d_lb, d_ulb)), n_
        out = S.ssl_step(lambda s, x, dl: U.unet_b_forward(s, x, True, domain_label=dl), st_s, st_t, {}, batch, n_classes=n_classes, branch=branch,
                         iter_num=iter_num, max_iterations=30000, lr=0.03, threshold=thr, domains=dom)
        assert torch.equal(out["loss"], loss.detach()), (out["loss"], loss)
        assert_same_state(ref_s, {k: v.detach() for k, v in st_s.items()})
        assert_same_state(ref_t, {k: v.detach() for k, v in st_t.items()})
    name = f"dsbnstep_b_{branch}_c{n_channels}_k{n_classes}_{hw}_b{B}_it{iter_num}_d{d_lb}{d_ulb}of{domains}"
    fx = {"loss": np_(loss), "terms": np.array([t.item() for t in terms]), "cw": np.float64(cw), "threshold": np.float64(thr),
          "d_lb": np.int64(d_lb), "d_ulb": np.int64(d_ulb), "domains": np.int64(domains)}
    for k, v in comp.items():
        fx["comp/" + k] = np_(v).astype(np.uint8)
    for k, v in out["logits"].items():
        fx["logits/" + k] = np_(v)
    for k, v in digest(ref_s.state_dict()).items():
        fx["student_after/" + k] = v
    for k, v in digest(ref_t.state_dict()).items():
        fx["teacher_after/" + k] = v
    fx["no_grad_params"] = np.array([k for k, v in out["grads"].items() if v is None])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print(f"{name}: loss={loss.item():.6f} ok")


def _reference_fft_functions():
    """train.py cannot be imported (argparse + dataset dispatch at import time), so the three FFT
    helpers are lifted out of its source with ``ast`` at generation time and executed unmodified; only
    ``random.uniform`` is replaced by a stub that replays the recorded ratios."""
    import ast
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    want = ("extract_amp_spectrum", "low_freq_mutate_np", "source_to_target_freq")
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want)

    class _Random:
        ratios = []

        @classmethod
        def uniform(cls, lo, hi):
            return cls.ratios.pop(0)

    ns = {"np": np, "random": _Random}
    exec(compile(code, os.path.join(REF, "train.py"), "exec"), ns)
    return ns, _Random


def case_fft_mix():
    """Frequency-domain style mix (train.py:158-207,628-636): the reference's own functions vs the oracle."""
    from oracle import fft_mix_ref as Fm
    ns, rnd = _reference_fft_functions()
    rng = np.random.RandomState(SEED)
    fx = {}
    for tag, (N, C, H, W, L) in {"c1_48x48_L0p1": (2, 1, 48, 48, 0.1), "c3_64x40_L0p05": (2, 3, 64, 40, 0.05),
                                 "c1_96x96_L0p01": (1, 1, 96, 96, 0.01), "c1_128x160_L0p03": (2, 1, 128, 160, 0.03)}.items():
        mix_img = torch.from_numpy(rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32))
        ulb = torch.from_numpy(rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32))
        ratios = rng.uniform(0, 1, N).tolist()
        rnd.ratios = list(ratios)
        out = []
        for i in range(N):                                            # train.py:629-633 verbatim
            amp_trg = ns["extract_amp_spectrum"]((ulb[i].cpu().numpy() + 1) * 127.5)
            img_freq = ns["source_to_target_freq"](((mix_img[i] + 1) * 127.5).cpu().numpy(), amp_trg, L=L, degree=1.0)
            img_freq = np.clip(img_freq, 0, 255).astype(np.float32)
            out.append(img_freq)
        ref = torch.tensor(np.array(out), dtype=torch.float32) / 127.5 - 1   # train.py:634-635
        got = Fm.move_transx(mix_img.numpy(), ulb.numpy(), ratios, L)
        assert np.array_equal(ref.numpy(), got), f"fft mix oracle differs from the reference ({tag})"
        fx[f"{tag}/mix_img"], fx[f"{tag}/ulb_x_w"] = mix_img.numpy(), ulb.numpy()
        fx[f"{tag}/ratio"], fx[f"{tag}/L"], fx[f"{tag}/out"] = np.asarray(ratios), np.asarray(L), ref.numpy()
    np.savez_compressed(os.path.join(OUT, "fft_mix.npz"), **fx)
    print("fft_mix: reference == oracle bit-for-bit on", len(fx) // 5, "cases")


def case_hardness():
    """train.py:705-718 with the reference's own utils/metrics.py functions vs the oracle restatement."""
    from oracle import hardness_ref as Hr
    from utils import metrics as ref_metrics               # reference
    rng = np.random.RandomState(SEED)
    fx = {}
    for tag, mode, fn, shape, hi in (("binary", "binary", ref_metrics.dice_coeff, (5, 24, 24), 2), ("2label", "2label", ref_metrics.dice_coeff_2label, (4, 2, 24, 24), 2),
                                     ("3label", "3label", ref_metrics.dice_coeff_3label, (4, 24, 24), 4)):
        stu = rng.randint(0, hi, shape).astype(np.int64)
        tea = np.where(rng.rand(*shape) < 0.8, stu, rng.randint(0, hi, shape)).astype(np.int64)
        stu[0] = 0                                          # a sample whose student map is empty ...
        tea[0] = 0                                          # ... and whose teacher map is empty too (dice := 0)
        tea[1] = 0
        arrs = fn(np.asarray(torch.from_numpy(stu).clone().cpu()), torch.from_numpy(tea).clone().cpu(), ret_arr=True)   # train.py:705
        n_part = len(arrs)
        tmp = arrs[0]
        for i in range(1, n_part):                           # train.py:706-710
            tmp += arrs[i]
        ref_h = 1 - tmp / n_part
        lq_idx, max_v = 0, -1
        for i in range(len(ref_h)):                          # train.py:714-718
            if ref_h[i] > max_v:
                max_v, lq_idx = ref_h[i], i
        h, lq, _ = Hr.hardness(stu, tea, mode)
        assert np.array_equal(h, ref_h) and lq == lq_idx, f"hardness oracle differs from the reference ({tag})"
        fx[f"{tag}/stu"], fx[f"{tag}/tea"], fx[f"{tag}/hardness"], fx[f"{tag}/lq_idx"] = stu.astype(np.uint8), tea.astype(np.uint8), ref_h, np.asarray(lq_idx)
    np.savez_compressed(os.path.join(OUT, "hardness.npz"), **fx)
    print("hardness: reference == oracle bit-for-bit on 3 modes")


def _reference_lines(path, start_pat, stop_pat, after=None, exclusive=False):
    """Source lines of a reference script from the first line containing ``start_pat`` (searched after the first line
    containing ``after``; ``exclusive``: from the line after it) up to, not including, the first later line containing
    ``stop_pat`` -- dedented, unmodified."""
    import textwrap
    lines = open(os.path.join(REF, path)).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if after in l) if after else 0
    a = next(i for i in range(i0, len(lines)) if start_pat in lines[i]) + (1 if exclusive else 0)
    b = next(i for i in range(a + 1, len(lines)) if stop_pat in lines[i])
    return textwrap.dedent("\n".join(lines[a:b])) + "\n", (a + 1, b)


def case_eval():
    """`test()` of the reference (train.py:279-302, train_mnms.py:549-552,...): its label-encoding and prediction lines,
    sliced out of the scripts and executed unmodified, plus its Dice functions (utils/metrics.py) vs oracle/eval_ref.py."""
    from oracle import eval_ref as Ev
    from oracle import hardness_ref as Hr
    from utils import metrics as ref_metrics               # reference
    enc_src, enc_at = _reference_lines("train.py", "mask = sample['label'].cuda()", "output = model(data)", after="def test(", exclusive=True)
    pred_src, pred_at = _reference_lines("train.py", "dice_loss(output, mask.unsqueeze(1), softmax=softmax", "mask, output = mask.cpu(), output.cpu()",
                                         after="def test(", exclusive=True)
    mnms_src, mnms_at = _reference_lines("train_mnms.py", "lb_mask = lb_y[:,...,0].eq(255).float()", "ulb_mask = ulb_y", after="for i_batch")
    dice_fn = {"fundus": ref_metrics.dice_coeff_2label, "prostate": ref_metrics.dice_coeff, "BUSI": ref_metrics.dice_coeff,
               "mnms": ref_metrics.dice_coeff_3label}           # train.py:220, train_mnms.py:212

    class _Args:
        dataset = None

    g = torch.Generator().manual_seed(SEED + 21)
    fx = {"src/encode_lines": np.asarray(enc_at), "src/predict_lines": np.asarray(pred_at), "src/mnms_lines": np.asarray(mnms_at)}
    for ds, k in (("fundus", 2), ("prostate", 2), ("BUSI", 2), ("mnms", 4)):
        B, H, W = 3, 24, 20
        if ds == "mnms":
            plane = torch.randint(0, 4, (B, H, W), generator=g)
            raw = torch.stack([(plane == c).to(torch.uint8) * 255 for c in (1, 2, 3)], dim=-1)         # RGB-coded label image
            ns = {"lb_y": raw.clone(), "torch": torch}
            exec(compile(mnms_src, "train_mnms.py", "exec"), ns)
            mask = ns["lb_mask"]
        else:
            raw = torch.tensor([0, 128, 255], dtype=torch.uint8)[torch.randint(0, 3, (B, H, W), generator=g)]
            a = _Args(); a.dataset = ds
            ns = {"mask": raw.clone(), "args": a, "torch": torch}
            exec(compile(enc_src, "train.py", "exec"), ns)
            mask = ns["mask"]
        output = torch.randn(B, k, H, W, generator=g) * 2
        if ds == "fundus":
            output[0] = -5.0                                  # an empty prediction for both structures
        if ds == "mnms":
            pred_label = torch.max(torch.softmax(output, dim=1), dim=1)[1]                                 # train_mnms.py:284
        else:
            a = _Args(); a.dataset = ds
            ns = {"output": output.clone(), "mask": mask.clone(), "args": a, "torch": torch}
            exec(compile(pred_src, "train.py", "exec"), ns)
            pred_label = ns["pred_label"]
        dice = dice_fn[ds](np.asarray(pred_label), mask)                                                 # train.py:302
        per_sample = dice_fn[ds](np.asarray(pred_label), mask, ret_arr=True)
        # oracle
        assert torch.equal(Ev.encode_labels(raw, ds), mask), f"eval oracle: label encoding differs ({ds})"
        assert torch.equal(Ev.predict(output, ds), pred_label), f"eval oracle: prediction differs ({ds})"
        got_dice, dc, jc = Ev.seg_metrics(pred_label.numpy(), mask.numpy(), ds)
        mode = {"prostate": "binary", "BUSI": "binary", "fundus": "2label", "mnms": "3label"}[ds]
        parts = Hr.dice_parts(pred_label.numpy(), mask.numpy(), mode)
        assert len(parts) == len(per_sample) and all(np.array_equal(a, b) for a, b in zip(parts, per_sample)), f"eval oracle: per-sample dice differs ({ds})"
        # batch mean: the reference does sum(list of Python floats) / len -- since Python 3.12 that sum is Neumaier-compensated,
        # so its last bit depends on the interpreter; the oracle (and the device kernel) add sequentially in double
        assert np.allclose(np.asarray(dice, dtype=np.float64), got_dice, rtol=4e-16, atol=0), f"eval oracle: dice differs ({ds})"
        fx[f"{ds}/raw"], fx[f"{ds}/mask"], fx[f"{ds}/output"] = raw.numpy(), mask.numpy(), output.numpy()
        fx[f"{ds}/pred_label"], fx[f"{ds}/dice"] = pred_label.numpy(), np.asarray(dice, dtype=np.float64)
        fx[f"{ds}/dice_per_sample"] = np.stack(per_sample)
        fx[f"{ds}/dc"], fx[f"{ds}/jc"] = dc, jc               # medpy restatement (not pinned: medpy is absent here)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **fx)
    print(f"eval: reference lines train.py:{enc_at[0]}-{enc_at[1]}, :{pred_at[0]}-{pred_at[1]}, train_mnms.py:{mnms_at[0]}-{mnms_at[1]} == oracle on 4 datasets")


def case_bank():
    """Confidence-bank bookkeeping (train.py:612-625, 720-743, 745-779) and ``obtain_all_cover_box`` (train.py:242-251): the
    reference's own lines, sliced out of train.py and executed unmodified on CPU tensors (``.cuda()`` is a no-op here, host RNG
    calls replay recorded draws), vs oracle/bank_ref.py."""
    import ast
    from oracle import bank_ref as Bk
    upd_src, upd_at = _reference_lines("train.py", "simple_ulb_idx = hardness < choice_th", "assert len(simple_ulb) == len(cor_pl)", after="for i_batch")
    pool_src, pool_at = _reference_lines("train.py", "if simple_ulb is None or len(simple_ulb) == 0:", "mix_img = cut_img[choice]", after="with amp_cm():")
    lq_src, lq_at = _reference_lines("train.py", "if args.dataset == 'fundus':", "logits_lq_s = model(lq_s)", after="new_choice = np.random.randint(0, len(lb_x_w))")
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    box_code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("obtain_all_cover_box", "obtain_cutmix_box"))
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        class _Args:
            increase, dataset = 1.0005, "prostate"

        class _NpRandom:
            draws = {}

            @classmethod
            def randint(cls, lo, hi, n=None):
                if lo == 0:
                    return np.asarray(cls.draws["r_lb"][:n], dtype=np.int64)
                nb = hi - lo
                return lo + np.floor(np.asarray(cls.draws["r_u"][:n], dtype=np.float64) * nb).astype(np.int64)

            @classmethod
            def permutation(cls, x):
                return np.asarray(x)[np.asarray(cls.draws["perm"])]

        class _Np:
            random = _NpRandom
            concatenate = staticmethod(np.concatenate)

        rng = np.random.RandomState(SEED + 31)
        g = torch.Generator().manual_seed(SEED + 31)
        fx = {"src/update_lines": np.asarray(upd_at), "src/pool_lines": np.asarray(pool_at), "src/lq_lines": np.asarray(lq_at)}
        for ds, C, k in (("prostate", 1, 2), ("fundus", 3, 2)):
            Bl = Bu = 4
            H = W = 16
            max_len = 6
            ref = dict(simple_ulb=None, cor_pl=None, cor_gt=None, cor_hardness=[], cor_dc=None, cor_mask=None, choice_th=0.1)
            st = Bk.new_state()
            # hardness scripts: nothing selected / some / all (overflow) / none again (threshold grows) ...
            script = [[0.5, 0.2, 0.3, 0.9], [0.05, 0.5, 0.02, 0.7], [0.01, 0.015, 0.5, 0.012], [0.3, 0.3, 0.3, 0.3], [0.001, 0.002, 0.003, 0.004],
                      [0.0005, 0.9, 0.0007, 0.0001], [0.5, 0.5, 0.5, 0.5], [0.00005, 0.00001, 0.5, 0.00002]]
            for step, hs in enumerate(script):
                hardness = np.asarray(hs, dtype=np.float64)
                ulb_x_w = torch.rand(Bu, C, H, W, generator=g) * 2 - 1
                lb_x_w = torch.rand(Bl, C, H, W, generator=g) * 2 - 1
                if ds == "fundus":
                    pseudo_label = torch.randint(0, 2, (Bu, 2, H, W), generator=g).float()
                    mask = torch.randint(0, 2, (Bu, 2, H, W), generator=g).float()
                    lb_mask = torch.randint(0, 2, (Bl, 2, H, W), generator=g).float()
                    lb_mask_shape = (Bl, 2, H, W)
                else:
                    pseudo_label = torch.randint(0, k, (Bu, H, W), generator=g)
                    mask = torch.randint(0, 2, (Bu, 1, H, W), generator=g).float()
                    lb_mask = torch.randint(0, k, (Bl, H, W), generator=g)
                    lb_mask_shape = (Bl, 1, H, W)
                # ---- pool + choice BEFORE the update (train.py:612-625), with the bank of the previous step
                n_bank = 0 if ref["simple_ulb"] is None else len(ref["simple_ulb"])
                draws = {"r_lb": rng.randint(0, Bl, Bu), "r_u": rng.uniform(0, 1, Bu), "perm": rng.permutation(Bu)}
                _NpRandom.draws = draws
                ns = dict(simple_ulb=ref["simple_ulb"], cor_pl=ref["cor_pl"], cor_mask=ref["cor_mask"], lb_x_w=lb_x_w, lb_mask=lb_mask, lb_mask_shape=lb_mask_shape,
                          ulb_x_s=ulb_x_w, torch=torch, np=_Np)
                exec(compile(pool_src, "train.py", "exec"), ns)
                ci, cl, cm = Bk.cut_pool(st, lb_x_w, lb_mask, lb_mask_shape)
                assert torch.equal(ns["cut_img"], ci) and torch.equal(ns["cut_label"], cl) and torch.equal(ns["cut_mask"], cm), f"bank oracle: pool differs ({ds}, step {step})"
                ch = Bk.draw_choice(n_bank, Bl, Bu, draws["r_lb"], draws["r_u"], draws["perm"])
                assert np.array_equal(np.asarray(ns["choice"]), ch), f"bank oracle: choice differs ({ds}, step {step})"
                # ---- update (train.py:745-779)
                a = _Args(); a.dataset = ds
                ns = dict(hardness=hardness, choice_th=ref["choice_th"], simple_ulb=ref["simple_ulb"], cor_pl=ref["cor_pl"], cor_gt=ref["cor_gt"], cor_hardness=ref["cor_hardness"],
                          cor_dc=ref["cor_dc"], cor_mask=ref["cor_mask"], ulb_x_w=ulb_x_w, pseudo_label=pseudo_label, ulb_mask=pseudo_label, ulb_dc=torch.arange(Bu), mask=mask,
                          max_len=max_len, args=a, torch=torch, np=np)
                exec(compile(upd_src, "train.py", "exec"), ns)
                for key in ("simple_ulb", "cor_pl", "cor_gt", "cor_hardness", "cor_dc", "cor_mask", "choice_th"):
                    ref[key] = ns[key]
                st = Bk.bank_update(st, hardness, ulb_x_w, pseudo_label, mask, max_len=max_len, increase=a.increase)
                assert st["choice_th"] == ref["choice_th"], (ds, step, st["choice_th"], ref["choice_th"])
                assert torch.equal(st["simple_ulb"], ref["simple_ulb"]) and torch.equal(st["cor_pl"], ref["cor_pl"]) and torch.equal(st["cor_mask"], ref["cor_mask"])
                assert np.array_equal(st["cor_hardness"], np.asarray(ref["cor_hardness"]))
                t = f"{ds}/{step}"
                fx[t + "/hardness"], fx[t + "/ulb_x_w"], fx[t + "/lb_x_w"] = hardness, np_(ulb_x_w), np_(lb_x_w)
                fx[t + "/pseudo_label"], fx[t + "/mask"], fx[t + "/lb_mask"] = np_(pseudo_label).astype(np.uint8), np_(mask).astype(np.uint8), np_(lb_mask).astype(np.uint8)
                fx[t + "/r_lb"], fx[t + "/r_u"], fx[t + "/perm"], fx[t + "/choice"] = draws["r_lb"], draws["r_u"], draws["perm"], np.asarray(ns_choice := ch)
                fx[t + "/n_bank_after"], fx[t + "/choice_th_after"] = np.int64(len(ref["simple_ulb"])), np.float64(ref["choice_th"])
                fx[t + "/bank_img"], fx[t + "/bank_pl"], fx[t + "/bank_mask"] = np_(ref["simple_ulb"]), np_(ref["cor_pl"]).astype(np.uint8), np_(ref["cor_mask"]).astype(np.uint8)
                fx[t + "/bank_hardness"] = np.asarray(ref["cor_hardness"], dtype=np.float64)
                # ---- low-quality sample CutMix with the previous pseudo label (train.py:720-738)
                lq_idx = int(np.argmax(hardness))
                lq_u, lq_pl, lq_mask = ulb_x_w[[lq_idx]].clone(), pseudo_label[[lq_idx]].clone(), mask[[lq_idx]].clone()
                new_choice = int(rng.randint(0, Bl))
                boxns = {"torch": torch, "np": np, "random": __import__("random")}
                exec(compile(box_code, "train.py", "exec"), boxns)
                ns = dict(args=a, lq_pl=lq_pl, lb_mask=lb_mask, new_choice=new_choice, obtain_all_cover_box=boxns["obtain_all_cover_box"], lq_u=lq_u, lb_x_w=lb_x_w, lq_mask=lq_mask, torch=torch)
                if bool((Bk.lq_region(lq_pl, lb_mask, new_choice, ds) != 0).any()):
                    exec(compile(lq_src, "train.py", "exec"), ns)
                    lq_s, pl_lq, m_lq, box = Bk.lq_compose(lq_u, lq_pl, lq_mask, lb_x_w, lb_mask, new_choice, ds)
                    assert torch.equal(ns["lq_s"], lq_s) and torch.equal(ns["pseudo_label_lq"], pl_lq) and torch.equal(ns["mask_lq"], m_lq), f"bank oracle: lq compose differs ({ds}, step {step})"
                    fx[t + "/lq_new_choice"], fx[t + "/lq_idx"], fx[t + "/lq_s"], fx[t + "/lq_box"] = np.int64(new_choice), np.int64(lq_idx), np_(lq_s), np_(box).astype(np.uint8)
        np.savez_compressed(os.path.join(OUT, "bank.npz"), **fx)
        print(f"bank: reference lines train.py:{upd_at[0]}-{upd_at[1]}, :{pool_at[0]}-{pool_at[1]}, :{lq_at[0]}-{lq_at[1]} == oracle on 2 datasets x 8 steps")
    finally:
        torch.Tensor.cuda = saved_cuda


def case_input():
    """Normalize_tf + ToTensor (dataloaders/custom_transforms.py:650-684, 728-753), the two reference classes lifted with ``ast``
    (the module itself imports matplotlib / skimage, which are absent here) and executed unmodified, vs oracle/input_ref.py."""
    import ast
    from oracle import input_ref as In
    src = open(os.path.join(REF, "dataloaders", "custom_transforms.py")).read()
    tree = ast.parse(src)
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("Normalize_tf", "ToTensor"))
    ns = {"np": np, "torch": torch, "GetBoundary": lambda: None}
    exec(compile(code, "custom_transforms.py", "exec"), ns)
    rng = np.random.RandomState(SEED + 41)
    fx = {}
    for tag, shape in (("rgb", (4, 24, 20, 3)), ("gray", (3, 16, 16))):
        imgs = rng.randint(0, 256, shape).astype(np.uint8)
        imgs.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)                      # every uint8 value occurs
        out = []
        for im in imgs:
            sample = {"image": im, "label": np.zeros(im.shape[:2], np.uint8), "strong_aug": im[::-1].copy()}
            sample = ns["ToTensor"]()(ns["Normalize_tf"]()(sample))
            out.append(sample["image"])
            assert torch.equal(sample["strong_aug"], In.normalize_to_tensor(im[::-1].copy()))
        ref = torch.stack(out)
        got = In.batch(imgs)
        assert torch.equal(ref, got), f"input oracle differs ({tag})"
        fx[tag + "/u8"], fx[tag + "/out"] = imgs, np_(ref)
    np.savez_compressed(os.path.join(OUT, "inputs.npz"), **fx)
    print("input: Normalize_tf + ToTensor (reference classes) == oracle on", len(fx) // 2, "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if len(sys.argv) > 1 and sys.argv[1] == "fft":
        case_fft_mix()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "hardness":
        case_hardness()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "eval":
        case_eval()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "input":
        case_input()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "bank":
        case_bank()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dsbn":
        case_unet_b_dsbn()
        case_step_dsbn()
        return
    case_fft_mix()
    case_hardness()
    case_bank()
    case_input()
    case_eval()
    case_losses()
    case_unet_a(1, 2, 32, 2)
    case_unet_a(3, 3, 32, 2)
    case_unet_b(3, 3, 32, 2)
    case_unet_b(1, 2, 48, 2)
    case_dsbn()
    case_unet_b_dsbn()
    case_step_dsbn()
    case_step("a", "softmax", 1, 2, 32, 2, iter_num=0, bank=0)
    case_step("a", "softmax", 1, 4, 32, 2, iter_num=15000, bank=2)
    case_step("b", "softmax", 3, 3, 32, 2, iter_num=3000, bank=0)
    case_step("b", "sigmoid", 3, 2, 32, 2, iter_num=3000, bank=2)


if __name__ == "__main__":
    main()
