"""Evaluation-side kernels: label encodings, prediction rule, Dice / dc / jc batch means --
integer / boolean work: BIT-EXACT against the oracle restatement; plus ``evaluate_batch`` against the eval-mode oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset", ["prostate", "BUSI", "fundus", "mnms"])
def test_encode_labels_bit_exact(dataset):
    from oracle import eval_ref as R
    from ustrun.evalops import encode_labels
    g = torch.Generator().manual_seed(1)
    vals = torch.tensor([0.0, 128.0, 255.0, 64.0, 200.0])
    shape = (3, 40, 56, 3) if dataset == "mnms" else (3, 40, 56)
    y = vals[torch.randint(0, 5, shape, generator=g)]
    ref = R.encode_labels(y.clone(), dataset)
    got = encode_labels(y.cuda(), dataset).cpu()
    assert got.dtype == torch.uint8 and tuple(got.shape) == tuple(ref.shape)
    assert torch.equal(got.long(), ref.long())
    with pytest.raises(ValueError):
        encode_labels(y.cuda()[0], dataset)


@pytest.mark.parametrize("dataset,C", [("prostate", 2), ("mnms", 4), ("fundus", 2)])
def test_predict_bit_exact_vs_torch_cuda(dataset, C):
    from oracle import eval_ref as R
    from ustrun.evalops import predict
    torch.manual_seed(2)
    out = torch.randn(4, C, 48, 64, device="cuda") * 3
    out[0, :, 0, :8] = 0.25                                   # exact ties: first index wins (softmax), sigmoid(0.25) >= 0.5
    out[1, 0, 1, :8] = 0.0                                    # sigmoid(0) == 0.5 -> ge is True
    ref = R.predict(out, dataset)                              # torch CUDA ops on the same device
    got = predict(out, "sigmoid" if dataset == "fundus" else "softmax")
    assert torch.equal(got.long(), ref.long())


@pytest.mark.parametrize("dataset,shape,hi", [("prostate", (5, 64, 64), 2), ("fundus", (4, 2, 48, 48), 2), ("mnms", (6, 72, 72), 4)])
def test_seg_metrics_bit_exact(dataset, shape, hi):
    from oracle import eval_ref as R
    from ustrun.evalops import seg_metrics
    rng = np.random.RandomState(4)
    pred = rng.randint(0, hi, shape).astype(np.uint8)
    tgt = np.where(rng.rand(*shape) < 0.85, pred, rng.randint(0, hi, shape)).astype(np.uint8)
    pred[0] = 0
    tgt[0] = 0                                                 # empty prediction and empty mask
    tgt[1] = 0
    dice, dc, jc = R.seg_metrics(pred, tgt, dataset)
    m = seg_metrics(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda(), dataset)
    assert np.array_equal(m["dice"].cpu().numpy(), dice)
    assert np.array_equal(m["dc"].cpu().numpy(), dc)
    assert np.array_equal(m["jc"].cpu().numpy(), jc)


def test_evaluate_batch_vs_eval_mode_oracle():
    """test() of the reference on one prostate-style batch: eval-mode UNet-A forward + loss_seg + metrics."""
    from networks.unet_model import UNet
    from oracle import eval_ref as R
    from oracle import ssl_step_ref as S
    from oracle import unet_ref as U
    from ustrun import engine as E
    from ustrun.evalops import evaluate_batch
    E.set_precision("fp32")
    try:
        torch.manual_seed(5)
        st = U.init_unet_a(1, 2, seed=5)
        for k_ in st:                                          # non-trivial running statistics
            if k_.endswith("running_mean"):
                st[k_] = torch.randn_like(st[k_]) * 0.1
            if k_.endswith("running_var"):
                st[k_] = torch.rand_like(st[k_]) + 0.5
        model = UNet(1, 2)
        model.load_state_dict(st)
        model = model.cuda().train()
        x = torch.rand(3, 1, 64, 64) * 2 - 1
        y = torch.tensor([0.0, 255.0])[torch.randint(0, 2, (3, 64, 64))]
        m = evaluate_batch(model, x.cuda(), y.cuda(), "prostate")
        assert model.training                                  # mode restored
        logits_ref = U.unet_a_forward({k_: v.clone() for k_, v in st.items()}, x, False)
        tgt = R.encode_labels(y, "prostate")
        loss_ref = S.masked_term(logits_ref, tgt, None, 2, "softmax")
        assert float((m["logits"].cpu() - logits_ref).norm() / logits_ref.norm()) < 1e-4
        assert abs(float(m["loss_seg"]) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
        dice, dc, jc = R.seg_metrics(m["pred"].cpu().numpy(), tgt.numpy(), "prostate")      # metrics of OUR prediction: bit-exact
        assert np.array_equal(m["dice"].cpu().numpy(), dice) and np.array_equal(m["dc"].cpu().numpy(), dc) and np.array_equal(m["jc"].cpu().numpy(), jc)
        agree = float((m["pred"].cpu().long() == R.predict(logits_ref, "prostate")).float().mean())
        assert agree > 0.999
    finally:
        E.set_precision("bf16")
