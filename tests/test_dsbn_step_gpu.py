"""DSBN as a whole network and as a whole SSL step (BASELINE.json configs[2]).

* ``networks.unet.UNet(norm='dsbn', num_domains=3)`` against the fixture produced by the REFERENCE's own
  ConvD / ConvU / UNet / DomainSpecificBatchNorm2d under the documented 6-line run-time patch (oracle/make_golden.py::_DsbnPatch).
* ``SSLTrainer.step`` with per-forward domain labels against the DSBN step fixture (reference modules + reference loss +
  torch.optim.SGD): losses, logits of all eight forwards, planes, student / teacher state after SGD + EMA -- including which
  domain's BatchNorm saw which forward (running statistics, num_batches_tracked) and that the third domain is untouched
  (no gradient => no SGD update, no weight decay; EMA still averages it: SURVEY F4).
* the UNet-A DSBN extension against the oracle.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_models_gpu import _check
from test_step_gpu import PLANES, check_state_after
from util import rel_err

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_b_dsbn_network(precision):
    from networks import unet as B
    from oracle import unet_ref as U
    dl = torch.tensor([1, 2])
    _check("unet_b_dsbn", precision, False, lambda: B.UNet(3, 2, norm="dsbn", num_domains=3),
           lambda: U.init_unet_b(3, 2, seed=1337, norm="dsbn", num_domains=3), lambda s, x: U.unet_b_forward(s, x, True, domain_label=dl),
           "unet_b_dsbn3_c3_k2_32.npz", 2, domain_label=dl)


def test_unet_b_dsbn_requires_label():
    from networks import unet as B
    m = B.UNet(3, 2, norm="dsbn", num_domains=3).cuda()
    with pytest.raises(TypeError):
        m(torch.zeros(1, 3, 32, 32, device="cuda"))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_a_dsbn_extension(precision):
    """No fixture (UNet-A has no DSBN upstream): the module against the oracle extension, which tests/test_oracle_golden.py holds
    equal to the plain network run with the selected domain's BatchNorm tensors."""
    from networks.unet_model import UNet
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import engine as E
    from util import load_state_into
    from utils.losses import MaskedCEDice
    E.set_precision(precision)
    try:
        st = U.init_unet_a(3, 2, seed=11, norm="dsbn", num_domains=3)
        g = torch.Generator().manual_seed(5)
        x = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
        tgt = torch.randint(0, 2, (2, 32, 32), generator=g)
        msk = (torch.rand(2, 1, 32, 32, generator=g) > 0.3).float()
        dl = torch.tensor([2, 2])
        ref = {k: v.clone().double() if v.is_floating_point() else v.clone() for k, v in st.items()}
        params, _ = U.split_state(ref)
        for p in params.values():
            p.requires_grad_(True)
        logits = U.unet_a_forward(ref, x.double(), True, domain_label=dl)
        loss = S.masked_term(logits, tgt, msk.double(), 2, "softmax")
        loss.backward()
        mod = load_state_into(UNet(3, 2, norm="dsbn", num_domains=3), st).cuda().train()
        out = mod(x.cuda(), domain_label=dl)
        l2 = MaskedCEDice(2)(out, tgt.cuda(), msk.cuda())
        l2.backward()
        tol = 1e-4 if precision == "fp32" else 3e-2
        assert rel_err(out, logits) < tol and abs(float(l2) - float(loss)) < tol * abs(float(loss))
        for n, p in mod.named_parameters():
            if ".bns." in n and ".bns.2." not in n:
                assert p.grad is None, n                 # the other domains' BatchNorms are not part of this forward
            else:
                assert p.grad is not None, n
        sd = mod.state_dict()
        assert int(sd["inc.double_conv.1.bns.2.num_batches_tracked"]) == 1 and int(sd["inc.double_conv.1.bns.0.num_batches_tracked"]) == 0
        if precision == "fp32":
            assert rel_err(mod.outc.conv.weight.grad, params["outc.conv.weight"].grad) < 1e-4
            assert rel_err(sd["down2.maxpool_conv.1.double_conv.4.bns.2.running_var"], ref["down2.maxpool_conv.1.double_conv.4.bns.2.running_var"]) < 1e-4
    finally:
        E.set_precision("bf16")


def _run_dsbn_step(precision, use_graph=False, steps=1):
    from networks import unet as B
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import engine as E
    from ustrun.step import SSLTrainer
    fx = np.load(os.path.join(GOLDEN, "dsbnstep_b_softmax_c3_k2_32_b2_it3000_d02of3.npz"))
    torch.manual_seed(1337)
    st_s, st_t = U.init_unet_b(3, 2, norm="dsbn", num_domains=3), U.init_unet_b(3, 2, norm="dsbn", num_domains=3)
    student, teacher = B.UNet(3, 2, norm="dsbn", num_domains=3), B.UNet(3, 2, norm="dsbn", num_domains=3)
    student.load_state_dict(st_s), teacher.load_state_dict(st_t)
    student, teacher = student.cuda().train(), teacher.cuda().train()
    for p in teacher.parameters():
        p.detach_()
    batch = S.synthetic_batch(3, 2, 32, 32, 2, 2, seed=1337)
    E.set_precision(precision)
    try:
        tr = SSLTrainer(student, teacher, n_classes=2, base_lr=0.03, max_iterations=30000, threshold=float(fx["threshold"]), use_graph=use_graph)
        tr.iter_num, tr.lr = 3000, 0.03
        outs = []
        for _ in range(steps):
            outs.append(tr.step({**{kk: v.cuda() for kk, v in batch.items()}, "domain_lb": int(fx["d_lb"]), "domain_ulb": int(fx["d_ulb"])}, keep_logits=True))
        torch.cuda.synchronize()
    finally:
        E.set_precision("bf16")
    return fx, outs, student, teacher, tr, (st_s, st_t)


def test_dsbn_step_fp32_matches_reference_fixture():
    fx, outs, student, teacher, tr, before = _run_dsbn_step("fp32")
    out = outs[0]
    assert abs(float(out["loss"]) - float(fx["loss"])) <= 1e-4 * abs(float(fx["loss"])), (float(out["loss"]), float(fx["loss"]))
    terms = [float(out[n]) for n in ("sup_loss", "unsup_loss_ul", "unsup_loss_lu", "unsup_loss_s")]
    assert np.allclose(terms, fx["terms"], rtol=1e-4, atol=1e-6), (terms, fx["terms"])
    for key in ("t1", "t2", "t3", "s0", "lb", "ul", "lu", "s"):
        assert rel_err(out["logits"][key].cpu(), torch.from_numpy(fx["logits/" + key])) < 1e-4, key
    for key in PLANES:
        got, want = out[key].cpu().numpy().astype(np.uint8), fx["comp/" + key].astype(np.uint8).reshape(out[key].shape)
        assert (got == want).mean() >= 0.999, (key, float((got == want).mean()))
    check_state_after(fx, student, teacher, before)
    # domain 1 is not in play: bit-identical student tensors (no SGD, no weight decay), zero forwards counted
    names = [n for n, _ in student.named_parameters()]
    untouched = set(fx["no_grad_params"].tolist())
    for n, p in student.named_parameters():
        if n in untouched:
            assert torch.equal(p.detach().cpu(), before[0][n]), n
    assert not any(tr.opt.has_grad[i] for i, n in enumerate(names) if n in untouched)
    sd = student.state_dict()
    assert [int(sd[f"convd1.bn1.bns.{d}.num_batches_tracked"]) for d in range(3)] == [2, 0, 3]


def test_dsbn_step_bf16_losses_within_tolerance():
    fx, outs, _, _, _, _ = _run_dsbn_step("bf16")
    assert abs(float(outs[0]["loss"]) - float(fx["loss"])) <= 1e-2 * abs(float(fx["loss"]))
