"""Whole-network parity against the oracle (CPU fp32 restatement pinned to the reference) and
against the committed golden fixtures generated from the reference itself.

Tolerances (BASELINE.json north_star): 1e-4 relative for the fp32 validation mode, 1e-2 relative
for bf16 (relative = ||a-b|| / ||b|| over the tensor)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import clone_state, load_state_into, rel_err

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # torch's fp32 reference convs must not use TF32
torch.backends.cuda.matmul.allow_tf32 = False

TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _loss_and_grads_oracle(fwd, st, x, tgt, msk, n_classes):
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    params, _ = U.split_state(st)
    for p in params.values():
        p.requires_grad_(True)
    logits = fwd(st, x)
    loss = S.masked_term(logits, tgt, msk, n_classes, "softmax")
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad for k, p in params.items()}


def _run_module(mod, x, tgt, msk, n_classes, **fkw):
    from utils.losses import MaskedCEDice
    logits = mod(x.cuda(), **fkw)
    loss = MaskedCEDice(n_classes)(logits, tgt.cuda(), msk.cuda())
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad for k, p in mod.named_parameters()}


def _check(name, precision, force_simt, make_mod, oracle_init, oracle_fwd, golden, n_classes, grad_tol_scale=3.0, **fkw):
    from ustrun import engine as E
    E.set_precision(precision)
    E.set_force_simt(force_simt)
    try:
        fx = np.load(os.path.join(GOLDEN, golden))
        x, tgt, msk = torch.from_numpy(fx["x"]), torch.from_numpy(fx["target"]), torch.from_numpy(fx["mask"])
        st = oracle_init()
        o_logits, o_loss, o_grads = _loss_and_grads_oracle(oracle_fwd, clone_state(st), x, tgt, msk, n_classes)
        # the oracle reproduces the reference fixture (bit-for-bit in the build container, see
        # tests/test_oracle_golden.py; another host CPU may pick different oneDNN kernels)
        assert np.allclose(o_logits.numpy(), fx["logits"], rtol=1e-4, atol=1e-5) and abs(float(o_loss) - float(fx["loss"])) < 1e-5
        mod = load_state_into(make_mod(), st).cuda().train()
        logits, loss, grads = _run_module(mod, x, tgt, msk, n_classes, **fkw)
        tol = TOL[precision]
        assert rel_err(logits, o_logits) < tol, f"{name}: logits {rel_err(logits, o_logits):.2e}"
        assert abs(float(loss) - float(o_loss)) / abs(float(o_loss)) < tol, f"{name}: loss"
        worst = 0.0
        for k, g in o_grads.items():
            if g is None:
                assert grads[k] is None, k
                continue
            assert grads[k] is not None, k
            if float(g.norm()) < 1e-6 * float(max(v.norm() for v in o_grads.values() if v is not None)):
                assert float(grads[k].float().norm()) < 1e-4 * float(max(v.norm() for v in o_grads.values() if v is not None)), k
                continue                      # conv bias in front of BatchNorm: mathematically zero
            worst = max(worst, rel_err(grads[k], g))
        assert worst < tol * grad_tol_scale, f"{name}: worst grad rel err {worst:.2e}"
        # running statistics
        sd = mod.state_dict()
        for k in sd:
            if k.endswith("running_mean") or k.endswith("running_var"):
                ref = fx["state_after/" + k]
                assert abs(float(sd[k].double().sum()) - ref[0]) <= tol * max(1.0, ref[1]), k
            if k.endswith("num_batches_tracked"):
                assert float(sd[k]) == fx["state_after/" + k][0], k
        return worst
    finally:
        E.set_precision("bf16")
        E.set_force_simt(False)


@pytest.mark.parametrize("precision,force_simt", [("fp32", True), ("bf16", True), ("bf16", False)])
@pytest.mark.parametrize("c,k", [(1, 2), (3, 3)])
def test_unet_a(precision, force_simt, c, k):
    from networks.unet_model import UNet
    from oracle import unet_ref as U
    _check(f"unet_a_{precision}", precision, force_simt, lambda: UNet(c, k), lambda: U.init_unet_a(c, k, seed=1337),
           lambda s, x: U.unet_a_forward(s, x, True), f"unet_a_c{c}_k{k}_32.npz", k)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("c,k,hw", [(3, 3, 32), (1, 2, 48)])
def test_unet_b(precision, c, k, hw):
    from networks import unet as B
    from oracle import unet_ref as U
    _check(f"unet_b_{precision}", precision, False, lambda: B.UNet(c, k), lambda: U.init_unet_b(c, k, seed=1337),
           lambda s, x: U.unet_b_forward(s, x, True), f"unet_b_c{c}_k{k}_{hw}.npz", k)


class _EncRec(torch.nn.Module):
    def __init__(self):
        super().__init__()
        from networks import unet as B
        self.enc = B.Encoder(c=3, norm="bn")
        self.dec = B.Rec_Decoder(num_classes=2, norm="dsbn", num_domains=3)

    def forward(self, x, domain_label=None):
        return self.dec(self.enc(x)[-1], domain_label=domain_label)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dsbn_encoder_rec_decoder(precision):
    """The only DSBN network constructible upstream: mixed-domain batch [2,2,0,1] ->
    the whole batch uses bns[2]; only that domain gets gradients / running-stat updates."""
    from oracle import unet_ref as U
    dl = torch.tensor([2, 2, 0, 1])

    def init():
        torch.manual_seed(1337)
        enc = U.init_unet_b(3, 2, norm="bn", decoder=False)
        dec = U.init_rec_decoder(num_classes=2, norm="dsbn", num_domains=3)
        return {**{"enc." + k: v for k, v in enc.items()}, **{"dec." + k: v for k, v in dec.items()}}

    def fwd(s, t):
        e = {k[4:]: v for k, v in s.items() if k.startswith("enc.")}
        d = {k[4:]: v for k, v in s.items() if k.startswith("dec.")}
        return U.rec_decoder_forward(d, U.unet_b_encoder(e, t, True)[-1], dl, True)

    _check("dsbn", precision, False, _EncRec, init, fwd, "dsbn_encrec.npz", 2, domain_label=dl)


def test_unet_a_eval_and_feature():
    from networks.unet_model import UNet
    from oracle import unet_ref as U
    from ustrun import engine as E
    E.set_precision("fp32")
    try:
        st = U.init_unet_a(1, 2, seed=1337)
        for k in st:
            if k.endswith("running_mean"):
                st[k].normal_(0, 0.1)
            if k.endswith("running_var"):
                st[k].uniform_(0.5, 1.5)
        x = torch.rand(1, 1, 32, 32) * 2 - 1
        ref_logits, ref_feat = U.unet_a_forward(clone_state(st), x, training=False, feature=True)
        mod = load_state_into(UNet(1, 2), st).cuda().eval()
        with torch.no_grad():
            logits, feat = mod(x.cuda(), feature=True)
        assert rel_err(logits, ref_logits) < 1e-4 and rel_err(feat, ref_feat) < 1e-4
        assert all(int(v) == 0 for k, v in mod.state_dict().items() if k.endswith("num_batches_tracked"))
    finally:
        E.set_precision("bf16")


def test_standalone_dsbn_module():
    from networks.dsbn import DomainSpecificBatchNorm2d
    from ustrun import engine as E
    E.set_precision("fp32")
    try:
        torch.manual_seed(5)
        m = DomainSpecificBatchNorm2d(16, 3).cuda()
        ref = torch.nn.BatchNorm2d(16).cuda()
        x = torch.randn(4, 16, 8, 8, device="cuda", requires_grad=True)
        x2 = x.detach().clone().requires_grad_()
        dl = torch.tensor([1, 0, 2, 2])
        y, out_dl = m(x, dl)
        yr = ref(x2)
        assert out_dl is dl and rel_err(y, yr) < 1e-5
        g = torch.randn_like(yr)
        y.backward(g)
        yr.backward(g)
        assert rel_err(x.grad, x2.grad) < 1e-4
        assert rel_err(m.bns[1].weight.grad, ref.weight.grad) < 1e-4 and m.bns[0].weight.grad is None
        assert [int(b.num_batches_tracked) for b in m.bns] == [0, 1, 0]
        assert rel_err(m.bns[1].running_var, ref.running_var) < 1e-5
    finally:
        E.set_precision("bf16")
