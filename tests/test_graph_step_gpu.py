"""The SSL step replayed as ONE CUDA graph (``SSLTrainer(use_graph=True)``) against the same step launched eagerly:
identical kernels on identical inputs, every per-step scalar (lr, EMA alpha, consistency weights) read from the device-side
hyper block -- so losses, planes and weights must agree bit for bit over several steps with changing inputs and schedules."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(kind, c, k, **kw):
    torch.manual_seed(1337)
    if kind == "a":
        from networks.unet_model import UNet
    else:
        from networks.unet import UNet
    s, t = UNet(c, k, **kw), UNet(c, k, **kw)
    t.load_state_dict(s.state_dict())
    for p in t.parameters():
        p.detach_()
    return s.cuda().train(), t.cuda().train()


@pytest.mark.parametrize("kind,c,k,hw,branch,hardness,mixdev", [("a", 1, 2, 64, "softmax", "binary", False), ("b", 3, 2, 64, "sigmoid", "2label", True),
                                                                 ("b", 3, 3, 48, "softmax", None, False)])
def test_graph_replay_equals_eager(kind, c, k, hw, branch, hardness, mixdev):
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    runs = []
    for use_graph in (False, True):
        s, t = _pair(kind, c, k)
        tr = SSLTrainer(s, t, n_classes=k, branch=branch, max_iterations=400, threshold=0.6, use_graph=use_graph, hardness_mode=hardness)
        tr.iter_num = 100                       # lr, alpha and the consistency weight all change from step to step here
        losses, planes = [], []
        for i in range(6):
            b = {kk: v.cuda() for kk, v in S.synthetic_batch(c, k, hw, hw, 2, 2, seed=50 + i, branch=branch).items()}
            if mixdev:
                del b["move_transx"]
                b["mix_ratio"] = torch.tensor([0.3, 0.7], dtype=torch.float64, device="cuda") * (i + 1) / 6
            out = tr.step(b, lq=b["ulb_w"][:1].contiguous())
            losses.append(out["loss"].detach().clone())
            planes.append(out["mask_w"].detach().clone())
            if hardness:
                planes.append(out["hardness"].detach().clone())
        torch.cuda.synchronize()
        runs.append((losses, planes, [p.detach().clone() for p in s.parameters()], [p.detach().clone() for p in t.parameters()],
                     [b.detach().clone() for b in s.buffers()], tr))
    (l0, p0, w0, t0, b0, tr0), (l1, p1, w1, t1, b1, tr1) = runs
    assert len(tr1._graphs) == 1 and tr1.launches_per_step > 100
    for a, b in zip(l0, l1):
        assert torch.equal(a, b), (float(a), float(b))
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)
    for a, b in zip(w0 + t0 + b0, w1 + t1 + b1):
        assert torch.equal(a, b)
    assert tr0.iter_num == tr1.iter_num == 106 and tr0.lr == tr1.lr


def test_graph_accepts_pinned_host_inputs():
    """The e2e path of bench.py: pinned host tensors are copied straight into the capture's static buffers."""
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    s, t = _pair("a", 1, 2)
    s2, t2 = _pair("a", 1, 2)
    tr = SSLTrainer(s, t, n_classes=2, threshold=0.6, use_graph=True)
    tr2 = SSLTrainer(s2, t2, n_classes=2, threshold=0.6)
    for i in range(5):
        host = S.synthetic_batch(1, 2, 64, 64, 2, 2, seed=7 + i)
        for kk in ("lb_mask", "cut_label", "cut_mask", "box"):
            host[kk] = host[kk].to(torch.uint8)
        host["choice"] = host["choice"].to(torch.int32)
        pinned = {kk: v.contiguous().pin_memory() for kk, v in host.items()}
        a = tr.step(pinned if i >= 2 else {kk: v.cuda() for kk, v in pinned.items()})
        b = tr2.step({kk: v.cuda() for kk, v in pinned.items()})
        assert torch.equal(a["loss"], b["loss"])
    for p, q in zip(s.parameters(), s2.parameters()):
        assert torch.equal(p, q)


def test_trainer_state_dict_roundtrip():
    """Checkpoint / resume (utils/util.py:259-297, train.py:542-548): momentum buffers, iter_num and lr survive."""
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    batches = [{kk: v.cuda() for kk, v in S.synthetic_batch(1, 2, 32, 32, 2, 2, seed=20 + i).items()} for i in range(4)]
    s, t = _pair("a", 1, 2)
    tr = SSLTrainer(s, t, n_classes=2, threshold=0.6)
    for b in batches[:2]:
        tr.step(b)
    ck = {"model": {k: v.clone() for k, v in s.state_dict().items()}, "ema": {k: v.clone() for k, v in t.state_dict().items()}, "trainer": tr.state_dict()}
    ref_sgd = torch.optim.SGD(s.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4)
    ref_sgd.load_state_dict(ck["trainer"]["optimizer"])          # the layout is torch.optim.SGD's own
    for b in batches[2:]:
        tr.step(b)
    s2, t2 = _pair("a", 1, 2)
    s2.load_state_dict(ck["model"]), t2.load_state_dict(ck["ema"])
    tr2 = SSLTrainer(s2, t2, n_classes=2, threshold=0.6)
    tr2.load_state_dict(ck["trainer"])
    for b in batches[2:]:
        tr2.step(b)
    assert tr2.iter_num == tr.iter_num and tr2.lr == tr.lr
    for p, q in zip(s.parameters(), s2.parameters()):
        assert torch.equal(p, q)


@pytest.mark.parametrize("kind,c,k,hw,branch,lanes,graph,dsbn", [("a", 1, 2, 64, "softmax", 2, False, False), ("a", 1, 2, 64, "softmax", 4, True, False),
                                                                  ("b", 3, 2, 64, "sigmoid", 4, True, False), ("b", 3, 2, 48, "softmax", 3, True, True)])
def test_multi_lane_step_is_bit_identical(kind, c, k, hw, branch, lanes, graph, dsbn):
    """lanes > 1 runs the independent forwards / loss branches of a step concurrently on several CUDA streams.  BatchNorm
    running statistics are logged per forward and applied in the reference's order (ustrun_bn_running_update), parameter
    gradients accumulate on the one weight-gradient stream in program order: every tensor of the step must equal the
    single-lane step bit for bit -- losses, planes, student / teacher weights, running statistics, num_batches_tracked."""
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    kw = dict(norm="dsbn", num_domains=3) if dsbn else {}
    runs = []
    for nl, ug in ((1, False), (lanes, graph)):
        s, t = _pair(kind, c, k, **kw)
        tr = SSLTrainer(s, t, n_classes=k, branch=branch, max_iterations=400, threshold=0.6, use_graph=ug, lanes=nl)
        tr.iter_num = 100
        losses, planes = [], []
        for i in range(6):
            b = {kk: v.cuda() for kk, v in S.synthetic_batch(c, k, hw, hw, 2, 2, seed=80 + i, branch=branch).items()}
            if dsbn:
                b.update(domain_lb=i % 2, domain_ulb=2)        # two signatures alternate: two graphs, two descriptor tables
            out = tr.step(b, lq=b["ulb_w"][:1].contiguous())
            losses.append(out["loss"].detach().clone())
            planes.append(out["mask_w"].detach().clone())
        torch.cuda.synchronize()
        runs.append((losses, planes, [p.detach().clone() for p in s.parameters()] + [p.detach().clone() for p in t.parameters()],
                     {n: b_.detach().clone() for n, b_ in list(s.named_buffers()) + [("t." + n, v) for n, v in t.named_buffers()]}, tr))
    (l0, p0, w0, b0, tr0), (l1, p1, w1, b1, tr1) = runs
    assert tr1.lanes == lanes and len(tr1._lane_streams) >= 2
    for a, b in zip(l0 + p0 + w0, l1 + p1 + w1):
        assert torch.equal(a, b)
    for n in b0:
        assert torch.equal(b0[n], b1[n]), n


@pytest.mark.parametrize("graph", [False, True])
def test_upload_prefetch_equals_direct_step(graph):
    """``trainer.upload(host batch)`` (copy stream, two staging sets) + ``step`` == ``step`` on device tensors, bit for bit."""
    from ustrun import synth as S
    from ustrun.step import SSLTrainer
    s, t = _pair("a", 1, 2)
    s2, t2 = _pair("a", 1, 2)
    tr = SSLTrainer(s, t, n_classes=2, threshold=0.6, use_graph=graph, lanes=2)
    tr2 = SSLTrainer(s2, t2, n_classes=2, threshold=0.6)

    def host(i):
        h = S.synthetic_batch(1, 2, 64, 64, 2, 2, seed=40 + i)
        for kk in ("lb_mask", "cut_label", "cut_mask", "box"):
            h[kk] = h[kk].to(torch.uint8)
        h["choice"] = h["choice"].to(torch.int32)
        return {kk: v.contiguous().pin_memory() for kk, v in h.items()}

    batches = [host(i) for i in range(6)]
    nxt = tr.upload(batches[0], lq=batches[0]["ulb_w"][:1])
    for i in range(6):
        cur = nxt
        a = tr.step(cur)
        if i + 1 < 6:
            nxt = tr.upload(batches[i + 1], lq=batches[i + 1]["ulb_w"][:1])
        b = tr2.step({kk: v.cuda() for kk, v in batches[i].items()}, lq=batches[i]["ulb_w"][:1].cuda())
        assert torch.equal(a["loss"], b["loss"]), i
    for p, q in zip(s.parameters(), s2.parameters()):
        assert torch.equal(p, q)
