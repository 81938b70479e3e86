"""CPU: the oracle restatement reproduces the golden fixtures that oracle/make_golden.py produced
by running the reference's own modules (bit-exact: same torch CPU ops in the same order)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ssl_step_ref as S
from oracle import unet_ref as U

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _digest(v):
    v = v.detach().double().flatten()
    return np.concatenate([[v.sum().item(), v.abs().sum().item()], v[:4].numpy()])


def _model_case(fixture, init, fwd, n_classes):
    fx = np.load(os.path.join(GOLDEN, fixture))
    st = init()
    params, _ = U.split_state(st)
    for p in params.values():
        p.requires_grad_(True)
    x, tgt, msk = torch.from_numpy(fx["x"]), torch.from_numpy(fx["target"]), torch.from_numpy(fx["mask"])
    logits = fwd(st, x)
    loss = S.masked_term(logits, tgt, msk, n_classes, "softmax")
    loss.backward()
    assert np.array_equal(logits.detach().numpy(), fx["logits"])
    assert float(loss) == float(fx["loss"])
    for k, p in params.items():
        key = "grad/" + k
        if key in fx.files:
            g = p.grad
            got = np.concatenate([[g.double().norm().item(), g.double().sum().item()], g.flatten()[:6].double().numpy()])
            np.testing.assert_allclose(got, fx[key], rtol=1e-12, atol=0)
        else:
            assert p.grad is None, k
    for k, v in st.items():
        np.testing.assert_allclose(_digest(v), fx["state_after/" + k], rtol=1e-12, atol=0, err_msg=k)


@pytest.mark.parametrize("c,k", [(1, 2), (3, 3)])
def test_unet_a_golden(c, k):
    _model_case(f"unet_a_c{c}_k{k}_32.npz", lambda: U.init_unet_a(c, k, seed=1337), lambda s, x: U.unet_a_forward(s, x, True), k)


@pytest.mark.parametrize("c,k,hw", [(3, 3, 32), (1, 2, 48)])
def test_unet_b_golden(c, k, hw):
    _model_case(f"unet_b_c{c}_k{k}_{hw}.npz", lambda: U.init_unet_b(c, k, seed=1337), lambda s, x: U.unet_b_forward(s, x, True), k)


def test_dsbn_golden():
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

    _model_case("dsbn_encrec.npz", init, fwd, 2)
    with pytest.raises(TypeError):      # unet.py:142-145: dsbn without a label
        U.rec_decoder_forward(U.init_rec_decoder(norm="dsbn", num_domains=2, seed=0), torch.zeros(1, 256, 2, 2), None)
    with pytest.raises(ValueError):     # dsbn.py:31-34
        U._dsbn({}, "x", torch.zeros(2, 3), [0], True)


def test_unet_b_dsbn_golden():
    """UNet-B with DSBN threaded through ConvD/ConvU.  The fixture is the output of the REFERENCE's own
    modules under the documented run-time patch of oracle/make_golden.py::_DsbnPatch (``normalization`` passes
    ``num_domains`` on, ``DSBN.forward`` takes its label from a context variable): 6 lines, none of them arithmetic."""
    dl = torch.tensor([1, 2])
    _model_case("unet_b_dsbn3_c3_k2_32.npz", lambda: U.init_unet_b(3, 2, seed=1337, norm="dsbn", num_domains=3),
                lambda s, x: U.unet_b_forward(s, x, True, domain_label=dl), 2)
    with pytest.raises(TypeError):      # the label is mandatory for a DSBN network
        U.unet_b_forward(U.init_unet_b(3, 2, seed=0, norm="dsbn", num_domains=2), torch.zeros(1, 3, 16, 16), True)


def step_domains(d_lb, d_ulb, B):
    """Same convention as oracle/make_golden.py::step_domains."""
    u, l = torch.full((B,), d_ulb, dtype=torch.long), torch.full((B,), d_lb, dtype=torch.long)
    return dict(t1=u, t2=u, t3=l, s0=u, lb=l, ul=u, lu=l, s=u, lq=u)


def test_dsbn_step_golden():
    path = os.path.join(GOLDEN, "dsbnstep_b_softmax_c3_k2_32_b2_it3000_d02of3.npz")
    fx = np.load(path)
    torch.manual_seed(1337)
    st_s, st_t = U.init_unet_b(3, 2, norm="dsbn", num_domains=3), U.init_unet_b(3, 2, norm="dsbn", num_domains=3)
    batch = S.synthetic_batch(3, 2, 32, 32, 2, 2, seed=1337)
    out = S.ssl_step(lambda s, x, dl: U.unet_b_forward(s, x, True, domain_label=dl), st_s, st_t, {}, batch, n_classes=2, iter_num=3000,
                     max_iterations=30000, lr=0.03, threshold=float(fx["threshold"]), domains=step_domains(int(fx["d_lb"]), int(fx["d_ulb"]), 2))
    assert float(out["loss"]) == float(fx["loss"])
    assert sorted(k for k, v in out["grads"].items() if v is None) == sorted(fx["no_grad_params"].tolist())
    assert all(".bns.1." in k for k in fx["no_grad_params"].tolist())          # domain 1 is not in play: no gradient, no SGD, no weight decay
    for k_, v in st_s.items():
        np.testing.assert_allclose(_digest(v), fx["student_after/" + k_], rtol=1e-12, atol=0, err_msg=k_)
    for k_, v in st_t.items():
        np.testing.assert_allclose(_digest(v), fx["teacher_after/" + k_], rtol=1e-12, atol=0, err_msg=k_)
    # running statistics: only the domains in play were updated
    for k_, v in st_s.items():
        if k_.endswith("num_batches_tracked") and ".bns." in k_:
            d = int(k_.split(".bns.")[1].split(".")[0])
            assert int(v) == {0: 2, 1: 0, 2: 3}[d], (k_, int(v))                 # labelled domain: lb, lu; unlabelled: s0, ul, s


def test_unet_a_dsbn_extension_equals_plain_bn_per_domain():
    """UNet-A has no DSBN upstream; the extension (init_unet_a(norm='dsbn')) must be exactly the plain network run with the
    selected domain's BatchNorm tensors."""
    st = U.init_unet_a(1, 2, seed=3, norm="dsbn", num_domains=2)
    plain = U.init_unet_a(1, 2, seed=3)
    for k in plain:
        if k.endswith(".weight") and plain[k].dim() == 4:
            assert torch.equal(plain[k], st[k])
    x = torch.rand(2, 1, 32, 32)
    sel = {k.replace(".bns.1", ""): v.clone() for k, v in st.items() if ".bns.0." not in k}
    assert list(sel.keys()) == list(plain.keys())
    assert torch.equal(U.unet_a_forward(st, x, True, domain_label=torch.tensor([1, 1])), U.unet_a_forward(sel, x, True))
    assert int(st["inc.double_conv.1.bns.0.num_batches_tracked"]) == 0 and int(st["inc.double_conv.1.bns.1.num_batches_tracked"]) == 1


def test_loss_scale_is_a_noop_in_fp32():
    torch.manual_seed(1337)
    a, b = U.init_unet_b(1, 2, seed=5), U.init_unet_b(1, 2, seed=5)
    batch = S.synthetic_batch(1, 2, 16, 16, 2, 2, seed=3)
    fwd = lambda s, x: U.unet_b_forward(s, x, True)
    o1 = S.ssl_step(fwd, a, {k: v.clone() for k, v in a.items()}, {}, batch, n_classes=2, threshold=0.6, iter_num=3000, update=False)
    o2 = S.ssl_step(fwd, b, {k: v.clone() for k, v in b.items()}, {}, batch, n_classes=2, threshold=0.6, iter_num=3000, update=False, loss_scale=1024.0)
    for k in o1["grads"]:
        if o1["grads"][k] is not None:
            assert torch.allclose(o1["grads"][k], o2["grads"][k], rtol=1e-5, atol=1e-9), k


def test_losses_golden():
    fx = np.load(os.path.join(GOLDEN, "losses.npz"))
    for C in (2, 3, 4):
        logits = torch.from_numpy(fx[f"softmax_C{C}/logits"]).requires_grad_()
        tgt, msk = torch.from_numpy(fx[f"softmax_C{C}/target"]), torch.from_numpy(fx[f"softmax_C{C}/mask"])
        for tag, m in (("m", msk), ("n", None)):
            loss = S.dice_loss_with_mask(logits, tgt.unsqueeze(1), C, mask=m, softmax=True)
            (g,) = torch.autograd.grad(loss, logits)
            assert float(loss) == float(fx[f"softmax_C{C}_{tag}/loss"]) and np.array_equal(g.numpy(), fx[f"softmax_C{C}_{tag}/grad"])
        z = S.dice_loss_with_mask(logits, tgt.unsqueeze(1), C, mask=torch.zeros_like(msk), softmax=True)
        assert float(z) == float(fx[f"softmax_C{C}_zero_mask/loss"])
        assert float(z) < 1.0     # F8: class 0 is never masked, so an all-zero mask is not "no loss"
    logits = torch.from_numpy(fx["sigmoid/logits"]).requires_grad_()
    tgt, msk = torch.from_numpy(fx["sigmoid/target"]), torch.from_numpy(fx["sigmoid/mask"])
    for tag, m in (("m", msk), ("n", None)):
        loss = S.dice_loss_with_mask(logits, tgt.unsqueeze(1), 2, mask=m, sigmoid=True, multi=True)
        assert float(loss) == float(fx[f"sigmoid_{tag}/loss"])
    for cur in (0, 1, 37.0, 199, 200, 500):
        assert S.sigmoid_rampup(cur, 200.0) == float(fx[f"rampup/{cur}"])
    with pytest.raises(AssertionError):
        S.dice_loss_with_mask(logits, tgt, 2, softmax=True, sigmoid=True)


def test_scalar_schedules():
    assert S.ema_alpha(0) == 0.0 and S.ema_alpha(1) == 0.5 and S.ema_alpha(10 ** 6) == 0.99     # train.py:91
    assert S.consistency_weight(0, 30000) == pytest.approx(np.exp(-5.0))
    assert S.consistency_weight(29999, 30000) == pytest.approx(float(np.exp(-5.0 * (1 - 199 / 200) ** 2)))
    assert S.poly_lr(0.03, 0, 30000) == 0.03


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "step_*.npz"))), ids=os.path.basename)
def test_step_golden(path):
    fx = np.load(path)
    name = os.path.basename(path)[:-4].split("_")
    model, branch = name[1], name[2]
    c, k, hw, B, it, bank = int(name[3][1:]), int(name[4][1:]), int(name[5]), int(name[6][1:]), int(name[7][2:]), int(name[8][4:])
    torch.manual_seed(1337)
    if model == "a":
        st_s, st_t = U.init_unet_a(c, k), U.init_unet_a(c, k)
        fwd = lambda s, x: U.unet_a_forward(s, x, True)
    else:
        st_s, st_t = U.init_unet_b(c, k), U.init_unet_b(c, k)
        fwd = lambda s, x: U.unet_b_forward(s, x, True)
    batch = S.synthetic_batch(c, k, hw, hw, B, B, seed=1337, branch=branch, bank=bank)
    out = S.ssl_step(fwd, st_s, st_t, {}, batch, n_classes=k, branch=branch, iter_num=it, max_iterations=30000, lr=0.03,
                     threshold=float(fx["threshold"]))
    assert float(out["loss"]) == float(fx["loss"])
    for key in ("pseudo_label", "mask", "pseudo_label_w", "mask_w", "pseudo_label_ul", "mask_ul", "pseudo_label_lu", "mask_lu"):
        assert np.array_equal(out[key].numpy().astype(np.uint8), fx["comp/" + key].astype(np.uint8)), key
    for k_, v in st_s.items():
        np.testing.assert_allclose(_digest(v), fx["student_after/" + k_], rtol=1e-12, atol=0, err_msg=k_)
    for k_, v in st_t.items():
        np.testing.assert_allclose(_digest(v), fx["teacher_after/" + k_], rtol=1e-12, atol=0, err_msg=k_)


@pytest.mark.skipif(not os.path.isdir("/root/reference/networks"), reason="reference checkout not present (GPU box)")
def test_oracle_live_against_reference():
    """In the build container the reference itself is importable: run it side by side."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, '/root/reference'); sys.path.insert(0, %r)\n"
            "import torch\n"
            "from networks import unet_model\n"
            "from oracle import unet_ref as U\n"
            "torch.manual_seed(7); m = unet_model.UNet(1, 2); m.train()\n"
            "st = U.init_unet_a(1, 2, seed=7)\n"
            "x = torch.rand(2, 1, 32, 32)\n"
            "assert torch.equal(m(x), U.unet_a_forward(st, x, True))\n"
            "print('live-ok')\n") % os.path.dirname(GOLDEN.rstrip('/')).rsplit('/tests', 1)[0]
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "live-ok" in r.stdout, r.stderr[-2000:]


def test_fft_mix_oracle_reproduces_reference_fixture():
    """oracle/fft_mix_ref.py against tests/golden/fft_mix.npz (outputs of the reference's own
    extract_amp_spectrum / low_freq_mutate_np / source_to_target_freq, train.py:158-207, lifted by make_golden)."""
    from oracle import fft_mix_ref as Fm
    fx = np.load(os.path.join(GOLDEN, "fft_mix.npz"))
    tags = sorted({k.split("/")[0] for k in fx.files})
    assert len(tags) == 4
    for tag in tags:
        got = Fm.move_transx(fx[f"{tag}/mix_img"], fx[f"{tag}/ulb_x_w"], fx[f"{tag}/ratio"].tolist(), float(fx[f"{tag}/L"]))
        assert got.dtype == np.float32
        # bit-for-bit in the build container; another BLAS/pocketfft build may differ in the last float64 bit
        assert np.allclose(got, fx[f"{tag}/out"], rtol=0, atol=2e-6), tag


def test_hardness_oracle_reproduces_reference_fixture():
    """oracle/hardness_ref.py against tests/golden/hardness.npz (outputs of the reference's utils/metrics.py
    dice_coeff / dice_coeff_2label / dice_coeff_3label through the train.py:705-718 bookkeeping)."""
    from oracle import hardness_ref as Hr
    fx = np.load(os.path.join(GOLDEN, "hardness.npz"))
    for mode in ("binary", "2label", "3label"):
        h, lq, _ = Hr.hardness(fx[f"{mode}/stu"], fx[f"{mode}/tea"], mode)
        assert np.array_equal(h, fx[f"{mode}/hardness"]) and lq == int(fx[f"{mode}/lq_idx"]), mode
        h1, lq1, _ = Hr.hardness(fx[f"{mode}/stu"], fx[f"{mode}/tea"], mode, first_epoch=True)
        assert np.all(h1 == 1) and lq1 == 0


def test_eval_oracle_reproduces_reference_fixture():
    """oracle/eval_ref.py against tests/golden/eval.npz: label encodings and predictions produced by the reference's own
    `test()` lines (train.py:279-286, :291-299, train_mnms.py:549-552, sliced out and executed by make_golden) and Dice values
    of its utils/metrics.py functions.  dc / jc are the medpy restatement (medpy is absent: recorded, not pinned)."""
    from oracle import eval_ref as Ev
    from oracle import hardness_ref as Hr
    fx = np.load(os.path.join(GOLDEN, "eval.npz"))
    for ds, mode in (("fundus", "2label"), ("prostate", "binary"), ("BUSI", "binary"), ("mnms", "3label")):
        raw, mask = torch.from_numpy(fx[f"{ds}/raw"]), torch.from_numpy(fx[f"{ds}/mask"])
        out = torch.from_numpy(fx[f"{ds}/output"])
        enc = Ev.encode_labels(raw, ds)
        assert enc.dtype == mask.dtype and torch.equal(enc, mask), ds
        pred = Ev.predict(out, ds)
        assert np.array_equal(pred.numpy(), fx[f"{ds}/pred_label"]), ds
        parts = Hr.dice_parts(pred.numpy(), mask.numpy(), mode)
        assert np.array_equal(np.stack(parts), fx[f"{ds}/dice_per_sample"]), ds
        dice, dc, jc = Ev.seg_metrics(pred.numpy(), mask.numpy(), ds)
        # batch mean: the reference's sum() of Python floats is Neumaier-compensated since Python 3.12 (last-bit differences)
        assert np.allclose(dice, fx[f"{ds}/dice"], rtol=4e-16, atol=0), ds
        assert np.array_equal(dc, fx[f"{ds}/dc"]) and np.array_equal(jc, fx[f"{ds}/jc"]), ds
    assert float(fx["fundus/dice_per_sample"][0, 0]) < 0.05          # the all-background prediction of sample 0


def test_bank_oracle_reproduces_reference_fixture():
    """oracle/bank_ref.py against tests/golden/bank.npz (outputs of the reference's own lines train.py:754-781, :612-626, :722-739
    and obtain_all_cover_box :242-251, executed by oracle/make_golden.py::case_bank)."""
    from oracle import bank_ref as Bk
    fx = np.load(os.path.join(GOLDEN, "bank.npz"))
    for ds in ("prostate", "fundus"):
        st = Bk.new_state()
        for step in range(8):
            p = f"{ds}/{step}"
            t = lambda k: torch.from_numpy(fx[p + "/" + k])
            n_bank = 0 if st["simple_ulb"] is None else len(st["simple_ulb"])
            ch = Bk.draw_choice(n_bank, 4, 4, fx[p + "/r_lb"], fx[p + "/r_u"], fx[p + "/perm"])
            assert np.array_equal(ch, fx[p + "/choice"])
            mask = t("mask").float()
            pl = t("pseudo_label").float() if ds == "fundus" else t("pseudo_label").long()
            st = Bk.bank_update(st, fx[p + "/hardness"], t("ulb_x_w"), pl, mask, max_len=6, increase=1.0005)
            assert len(st["simple_ulb"]) == int(fx[p + "/n_bank_after"]) and st["choice_th"] == float(fx[p + "/choice_th_after"])
            assert np.array_equal(st["simple_ulb"].numpy(), fx[p + "/bank_img"]) and np.array_equal(st["cor_hardness"], fx[p + "/bank_hardness"])
            assert np.array_equal(st["cor_pl"].numpy().astype(np.uint8), fx[p + "/bank_pl"]) and np.array_equal(st["cor_mask"].numpy().astype(np.uint8), fx[p + "/bank_mask"])
            if (p + "/lq_s") in fx.files:
                i, nc = int(fx[p + "/lq_idx"]), int(fx[p + "/lq_new_choice"])
                lbm = t("lb_mask").float() if ds == "fundus" else t("lb_mask").long()
                lq_s, _, _, box = Bk.lq_compose(t("ulb_x_w")[[i]], pl[[i]], mask[[i]], t("lb_x_w"), lbm, nc, ds)
                assert np.array_equal(lq_s.numpy(), fx[p + "/lq_s"]) and np.array_equal(box.numpy().astype(np.uint8), fx[p + "/lq_box"])
        assert float(fx[f"{ds}/7/choice_th_after"]) < 0.1 and int(fx[f"{ds}/2/n_bank_after"]) >= 3


def test_input_oracle_reproduces_reference_fixture():
    """oracle/input_ref.py against tests/golden/inputs.npz (Normalize_tf + ToTensor of the reference, executed by make_golden)."""
    from oracle import input_ref as In
    fx = np.load(os.path.join(GOLDEN, "inputs.npz"))
    for tag in ("rgb", "gray"):
        assert np.array_equal(In.batch(fx[tag + "/u8"]).numpy(), fx[tag + "/out"]), tag
    assert float(In.normalize_to_tensor(np.full((2, 2), 255, np.uint8)).max()) == 1.0 and float(In.normalize_to_tensor(np.zeros((2, 2), np.uint8)).min()) == -1.0
