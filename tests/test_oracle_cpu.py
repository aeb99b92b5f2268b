"""Pins the CPU oracle (oracle/mcan_oracle.py) against the golden fixtures that
oracle/make_golden.py produced from the UNMODIFIED reference, and -- when /root/reference is
mounted (build container only) -- against the live reference modules."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mcan_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {
    "tiny_dense": (orc.TINY, "none"),
    "tiny_prefix": (orc.TINY, "prefix"),
    "tiny_random": (orc.TINY, "random"),
    "tiny_d128": (dict(orc.TINY, hidden_size=256, multi_head=2, flat_glimpses=1), "random"),
}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden_whole_net(name):
    cfgd, _ = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    batch, regions, tokens, token_size, answer_size, wseed, bseed = [int(v) for v in g["meta"]]
    cfg = orc.Cfg(dropout_rate=0.0, **cfgd)
    sd = orc.synth_state_dict(cfg, token_size, answer_size, seed=wseed, dtype=torch.float64)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    v = torch.from_numpy(g["img_feat"])
    q = torch.from_numpy(g["ques_ix"])
    ans = torch.from_numpy(g["ans"])
    probs, v_out, v_mask, v_w, q_out, q_mask, q_w, a = orc.net_forward(p, v, q, cfg)
    assert np.array_equal(v_mask.numpy(), g["v_mask"]) and np.array_equal(q_mask.numpy(), g["q_mask"])
    for got, key in ((probs, "probs"), (v_out, "v"), (q_out, "q"), (v_w, "v_w"), (q_w, "q_w"), (a, "a")):
        assert _rel(got.detach().numpy(), g[key]) < 1e-10, key
    assert (probs.argmax(1).numpy() == g["probs"].argmax(1)).all()
    loss = orc.bce_sum(probs, ans)
    assert abs(loss.item() - float(g["loss"])) < 1e-8 * abs(float(g["loss"]))
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    assert names == [n for n, _ in orc.param_shapes(cfg, token_size, answer_size)]
    for n, dig in zip(names, g["grad_digests"]):
        mine = orc.grad_digest(p[n].grad)
        assert np.abs(mine - dig).max() < 1e-8 * max(1.0, np.abs(dig).max()), n


def test_oracle_matches_reference_golden_modules():
    g = np.load(os.path.join(GOLD, "modules_tiny.npz"))
    cfg = orc.Cfg(dropout_rate=0.0, **orc.TINY)
    x0, y0 = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    x_mask, y_mask = torch.from_numpy(g["x_mask"]), torch.from_numpy(g["y_mask"])

    def run(tag, fn):
        names = [str(n) for n in g[tag + "_pnames"]]
        params = orc.seeded_params(names, g[tag + "_pshapes"], orc.MODULE_SEEDS[tag])
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        x = x0.clone().requires_grad_(True)
        y = y0.clone().requires_grad_(True)
        res = fn(p, x, y)
        out = res[0] if isinstance(res, tuple) else res
        assert _rel(out.detach().numpy(), g[tag + "_out"]) < 1e-10, tag
        if isinstance(res, tuple):
            assert _rel(res[1].detach().numpy(), g[tag + "_out2"]) < 1e-10, tag
        out.backward(torch.from_numpy(g[tag + "_gout"]))
        if tag + "_dx" in g.files:
            assert _rel(x.grad.numpy(), g[tag + "_dx"]) < 1e-9, tag
        if tag + "_dy" in g.files:
            assert _rel(y.grad.numpy(), g[tag + "_dy"]) < 1e-9, tag
        for n, dig in zip(names, g[tag + "_gdigests"]):
            mine = orc.grad_digest(p[n].grad)
            assert np.abs(mine - dig).max() < 1e-8 * max(1.0, np.abs(dig).max()), (tag, n)

    run("ln", lambda p, x, y: orc.layer_norm(x, p["a_2"], p["b_2"]))
    run("mhatt_self", lambda p, x, y: orc.mhatt(p, "", x, x, x, x_mask, cfg))
    run("mhatt_guided", lambda p, x, y: orc.mhatt(p, "", y, y, x, y_mask, cfg))
    run("sa", lambda p, x, y: orc.sa(p, "", x, x_mask, cfg))
    run("sga", lambda p, x, y: orc.sga(p, "", x, y, x_mask, y_mask, cfg))
    run("attflat", lambda p, x, y: orc.attflat(p, "", x, x_mask, cfg))


def test_closed_form_backwards_match_autograd():
    """The formulas the CUDA backward kernels implement == autograd of the forward restatement."""
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((5, 7, 32))).requires_grad_(True)
    a2 = torch.from_numpy(1 + 0.1 * rs.standard_normal(32)).requires_grad_(True)
    b2 = torch.from_numpy(rs.standard_normal(32)).requires_grad_(True)
    dy = torch.from_numpy(rs.standard_normal((5, 7, 32)))
    orc.layer_norm(x, a2, b2).backward(dy)
    dx, da, db = orc.layer_norm_backward(dy, x.detach(), a2.detach())
    assert _rel(dx.numpy(), x.grad.numpy()) < 1e-12
    assert _rel(da.numpy(), a2.grad.numpy()) < 1e-12 and _rel(db.numpy(), b2.grad.numpy()) < 1e-12

    q = torch.from_numpy(rs.standard_normal((2, 3, 6, 8))).requires_grad_(True)
    k = torch.from_numpy(rs.standard_normal((2, 3, 5, 8))).requires_grad_(True)
    v = torch.from_numpy(rs.standard_normal((2, 3, 5, 8))).requires_grad_(True)
    mask = torch.from_numpy(rs.uniform(size=(2, 1, 1, 5)) < 0.4)
    mask[1] = True
    do = torch.from_numpy(rs.standard_normal((2, 3, 6, 8)))
    orc.attention(v, k, q, mask).backward(do)
    dv, dk, dq = orc.attention_backward(do, v.detach(), k.detach(), q.detach(), mask)
    assert _rel(dv.numpy(), v.grad.numpy()) < 1e-12
    assert _rel(dk.numpy(), k.grad.numpy()) < 1e-12 and _rel(dq.numpy(), q.grad.numpy()) < 1e-12


def test_reference_semantics_edge_cases():
    # LayerNorm is NOT F.layer_norm (unbiased std, eps on std)
    x = torch.randn(4, 512, dtype=torch.float64)
    mine = orc.layer_norm(x, torch.ones(512, dtype=torch.float64), torch.zeros(512, dtype=torch.float64))
    assert (mine - torch.nn.functional.layer_norm(x, (512,))).abs().max() > 1e-4
    # constant row -> b_2
    c = torch.full((2, 16), 3.0, dtype=torch.float64)
    b2 = torch.arange(16, dtype=torch.float64)
    assert torch.equal(orc.layer_norm(c, torch.ones(16, dtype=torch.float64), b2), b2.expand(2, 16))
    # fully masked keys -> uniform softmax, not NaN
    q = torch.randn(1, 1, 3, 4, dtype=torch.float64)
    k = torch.randn(1, 1, 5, 4, dtype=torch.float64)
    v = torch.randn(1, 1, 5, 4, dtype=torch.float64)
    out = orc.attention(v, k, q, torch.ones(1, 1, 1, 5, dtype=torch.bool))
    assert torch.allclose(out, v.mean(2, keepdim=True).expand_as(out))
    # warm-up schedule (optim.py:36-49)
    got = [orc.warmup_rate(s, 1e-4, 100, 10) for s in (1, 10, 11, 20, 21, 30, 31)]
    assert np.allclose(got, [2.5e-5, 2.5e-5, 5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], rtol=1e-12, atol=0)


def test_synth_is_deterministic_and_matches_state_dict_contract():
    cfg = orc.Cfg(**orc.SMALL)
    shapes = orc.param_shapes(cfg, 20000, 3129)
    assert len(shapes) == 275                       # Net small has 275 tensors
    assert sum(int(np.prod(s)) for _, s in shapes) == 55512507
    cfgl = orc.Cfg(**orc.LARGE)
    assert sum(int(np.prod(s)) for _, s in orc.param_shapes(cfgl, 20000, 3129)) == 201551291
    a = orc.synth_state_dict(orc.Cfg(**orc.TINY), 50, 24, seed=3)
    b = orc.synth_state_dict(orc.Cfg(**orc.TINY), 50, 24, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.skipif(not os.path.isdir("/root/reference/core/model"), reason="reference not mounted")
def test_oracle_matches_live_reference_small_config():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden
    ref_net, _, _ = make_golden.import_reference()
    cfg = orc.Cfg(dropout_rate=0.0, **orc.SMALL)
    torch.manual_seed(0)
    net = ref_net.Net(cfg, None, 200, 3129).eval()      # reference's own default init
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    v, q, ans = orc.synth_batch(cfg, 2, 20, 14, 200, 3129, seed=5, ragged="prefix")
    with torch.no_grad():
        ref = net(v, q)
        mine = orc.net_forward(sd, v, q, cfg)
    assert (ref[0] - mine[0]).abs().max() < 1e-5
    assert (ref[1] - mine[1]).abs().max() < 1e-4 and (ref[4] - mine[4]).abs().max() < 1e-4


@pytest.mark.skipif(not os.path.isdir("/root/reference/core/model"), reason="reference not mounted")
def test_oracle_classifier_matches_live_reference():
    """ClassifierNet (reference net.py:138-184: SA-only stack over the image): outputs and
    parameter gradients of the oracle restatement vs the unmodified reference module."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden
    ref_net, _, _ = make_golden.import_reference()
    cfg = orc.Cfg(dropout_rate=0.0, **dict(orc.TINY, layer=2))
    torch.manual_seed(1)
    net = ref_net.ClassifierNet(cfg, 24).double().train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert set(sd) == {n for n, _ in orc.param_shapes(cfg, 50, 24, classifier=True)}
    v, _, ans = orc.synth_batch(cfg, 3, 12, 7, 50, 24, seed=6, ragged="random")
    v, ans = v.double(), ans.double()
    ref = net(v)
    orc.bce_sum(ref[0], ans).backward()
    p = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
    mine = orc.classifier_forward(p, v, cfg)
    orc.bce_sum(mine[0], ans).backward()
    for a, b in zip(ref, mine):
        assert a.shape == b.shape and (a.double() - b.double()).abs().max() < 1e-10
    gmax = max(float(q.grad.abs().max()) for q in net.parameters() if q.grad is not None)
    for n, q in net.named_parameters():
        if q.grad is None:            # attflat_lang is part of the state_dict but unused (net.py:150)
            assert n.startswith("attflat_lang.") and p[n].grad is None
            continue
        # (the key biases have a mathematically zero gradient -- softmax is shift invariant -- so the
        # error is measured against the largest gradient entry of the model, not per tensor)
        assert np.abs(p[n].grad.numpy() - q.grad.numpy()).max() < 1e-10 * gmax, n


def test_oracle_classifier_matches_reference_golden():
    """ClassifierNet fixture produced by the unmodified reference (oracle/make_golden.py classifier): runs
    wherever the repository is, also on the GPU box where /root/reference does not exist."""
    g = np.load(os.path.join(GOLD, "classifier_tiny.npz"))
    batch, regions, answer_size, wseed, bseed = [int(v) for v in g["meta"]]
    cfg = orc.Cfg(dropout_rate=0.0, **dict(orc.TINY, layer=2))
    sd = orc.synth_state_dict(cfg, 50, answer_size, seed=wseed, dtype=torch.float64, classifier=True)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    probs, v_out, v_mask, v_w, a = orc.classifier_forward(p, torch.from_numpy(g["img_feat"]), cfg)
    assert np.array_equal(v_mask.numpy(), g["v_mask"])
    for got, key in ((probs, "probs"), (v_out, "v"), (v_w, "v_w"), (a, "a")):
        assert _rel(got.detach().numpy(), g[key]) < 1e-10, key
    loss = orc.bce_sum(probs, torch.from_numpy(g["ans"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-8 * abs(float(g["loss"]))
    loss.backward()
    for n, dig in zip([str(n) for n in g["grad_names"]], g["grad_digests"]):
        mine = orc.grad_digest(p[n].grad)
        assert np.abs(mine - dig).max() < 1e-8 * max(1.0, np.abs(dig).max()), n
    unused = [str(n) for n in g["no_grad_names"]]
    assert unused and all(n.startswith("attflat_lang.") and p[n].grad is None for n in unused)
