"""Host-side logic that needs no GPU: the GEMM launch planner (tile / cluster / split-K / tail
splitting choice, include/mcan_b200.h: mcan_gemm_plan) and the protocol by which an optimiser
keeps the bf16 operand copies of blocks.LinearParams current."""
import pytest
import torch

SMS = 148        # B200


@pytest.fixture(scope="module")
def ops():
    from mcan_vqa_b200 import ops as o
    return o


def _tiles(pl):
    return pl["m_tiles"] * pl["n_tiles"]


@pytest.mark.parametrize("m,n,k,acc", [
    (6400, 4096, 1024, False), (6400, 1024, 4096, False), (6400, 3072, 1024, False), (6400, 1024, 1024, False),
    (896, 1024, 4096, False), (896, 4096, 1024, False), (896, 1024, 1024, False), (64, 3129, 2048, False),
    (1024, 1024, 6400, True), (4096, 1024, 6400, True), (1024, 1024, 896, True), (896, 1024, 4096, True),
    (100, 512, 512, False), (42, 136, 64, False), (1, 8, 8, False),
])
def test_plan_invariants(ops, m, n, k, acc):
    pl = ops.gemm_plan(m, n, k, accumulate=acc, sms=SMS)
    assert pl["block_n"] in (64, 128, 256) and pl["cluster"] in (1, 2)       # clusters of 4 are never auto-selected
    assert pl["m_tiles"] * 128 * pl["cluster"] >= m and pl["n_tiles"] * pl["block_n"] >= n
    assert (pl["m_tiles"] - 1) * 128 * pl["cluster"] < m and (pl["n_tiles"] - 1) * pl["block_n"] < n
    kblocks = (k + 63) // 64
    assert 1 <= pl["splits"] <= kblocks and (acc or pl["splits"] == 1)
    tiles = _tiles(pl)
    assert 0 < pl["full_tiles"] <= tiles
    if pl["splits"] > 1:
        assert pl["full_tiles"] == tiles and pl["units"] == tiles * pl["splits"]
    else:
        assert pl["units"] == pl["full_tiles"] + 2 * (tiles - pl["full_tiles"])
    if pl["full_tiles"] < tiles:          # tail splitting: whole waves at full width, the rest fits one half-wave
        slots = SMS // pl["cluster"]
        assert pl["block_n"] == 256 and n % 256 == 0 and pl["full_tiles"] % slots == 0
        assert 2 * (tiles - pl["full_tiles"]) <= slots


def test_plan_of_the_mcan_large_shapes(ops):
    """Regression guard for the choices the measurements in DESIGN.md section 4 were taken with."""
    ffn1 = ops.gemm_plan(6400, 4096, 1024, sms=SMS)          # 400 tiles on 74 pairs: 5 whole waves + 30 tiles split
    assert (ffn1["block_n"], ffn1["cluster"], ffn1["full_tiles"], ffn1["units"]) == (256, 2, 370, 430)
    merge = ops.gemm_plan(6400, 1024, 1024, sms=SMS)         # 100 tiles: 74 full + 26 x 2 halves
    assert (merge["block_n"], merge["cluster"], merge["full_tiles"], merge["units"]) == (256, 2, 74, 126)
    enc = ops.gemm_plan(896, 1024, 1024, sms=SMS)            # short M, short K: 128 x 64 single-CTA tiles
    assert (enc["block_n"], enc["cluster"]) == (64, 1) and enc["units"] == 7 * 16
    enc_sk = ops.gemm_plan(896, 1024, 4096, accumulate=True, sms=SMS)    # split-K with the linear fused epilogue:
    # short M -> single-CTA 128 x 128 tiles (profiles/r02_tile_sweep.txt), K splits fill the SMs
    assert enc_sk["block_n"] == 128 and enc_sk["cluster"] == 1 and enc_sk["splits"] >= 2
    assert _tiles(enc_sk) == 7 * 8 and _tiles(enc_sk) * enc_sk["splits"] <= 2 * SMS
    wgrad = ops.gemm_plan(1024, 1024, 6400, accumulate=True, sms=SMS)    # deep split-K: the 256 x 256 pair tile
    assert (wgrad["block_n"], wgrad["cluster"], wgrad["splits"]) == (256, 2, 4)
    head = ops.gemm_plan(64, 3129, 2048, sms=SMS)
    assert head["cluster"] == 1 and head["n_tiles"] * head["block_n"] >= 3129


def test_plan_honours_forced_configuration_and_sm_limit(ops):
    pl = ops.gemm_plan(6400, 4096, 1024, block_n=256, cta_group=4, sms=SMS)      # two pairs per cluster: 512-row super-tiles
    assert pl["cluster"] == 4 and pl["m_tiles"] == 13 and pl["full_tiles"] == _tiles(pl)
    pl = ops.gemm_plan(6400, 1024, 1024, cta_group=2, sms=SMS)
    assert pl["cluster"] == 2 and pl["block_n"] in (128, 256)
    few = ops.gemm_plan(6400, 4096, 1024, sms=40)             # persistent grids shrink with mcan_set_sm_limit
    assert few["units"] >= _tiles(few)
    from mcan_vqa_b200 import capi
    with pytest.raises(capi.McanError):
        ops.gemm_plan(6400, 4096, 1024, block_n=64, cta_group=2, sms=SMS)


class _FakeOptimizer(object):
    epoch = 0


def test_linear_params_shadow_protocol():
    """blocks.LinearParams: first use casts; once an optimiser manages the copy nothing is re-cast
    (even on the forced per-forward refresh) until a master changes behind its back; the low-order
    halves of the split-precision mode are refreshed once per optimiser epoch."""
    from mcan_vqa_b200.blocks import LinearParams
    lin1, lin2 = torch.nn.Linear(16, 8), torch.nn.Linear(16, 24)
    lp = LinearParams([(lin1.weight, lin1.bias), (lin2.weight, lin2.bias)])
    items = lp.shadow_items()
    assert [tuple(d.shape) for d, _ in items] == [(8, 16), (8,), (24, 16), (24,)]
    assert items[0][0].dtype == torch.bfloat16 and items[1][0].dtype == torch.float32
    assert items[0][0].data_ptr() == lp.w.data_ptr() and items[2][0].data_ptr() == lp.w[8:].data_ptr()
    assert len(lp.pending()) == 4                 # first use: everything
    assert lp.pending() == []                     # current
    assert len(lp.pending(force=True)) == 4       # unmanaged: a training forward re-casts
    opt = _FakeOptimizer()
    lp.managed = opt
    lp.pending()
    assert lp.pending(force=True) == []           # managed: the optimiser keeps the copy current
    with torch.no_grad():
        lin1.weight.add_(1.0)                     # e.g. load_state_dict: version counter moves
    assert len(lp.pending()) == 4
    assert lp.pending(force=True) == []
    # split precision: lo halves are created on demand and refreshed once per optimiser epoch
    first = lp.pending(need_lo=True)
    assert len(first) == 4 and len(first[0]) == 3
    assert lp.pending(need_lo=True) == []
    opt.epoch += 1
    assert len(lp.pending(need_lo=True)) == 4
    assert lp.pending(need_lo=True) == []
    # a weight whose row length is not a multiple of 8 has a padded copy: not eligible for shadowing
    odd = torch.nn.Linear(12, 4)
    assert LinearParams([(odd.weight, odd.bias)]).shadow_items() is None


def test_attflat_fc_gradient_bf16_noise_floor():
    """Why the first layer of the AttFlat MLP gets its own gradient tolerance (tests/test_reference_gpu.py):
    in an otherwise fp64 AttFlat (reference net.py:38-55), rounding ONLY the input of that GEMM to bf16 -- what
    any bf16 tensor-core GEMM does -- moves the gradient of mlp.fc.linear.weight by percent, ten times more than
    the 0.4 % rounding step: the per-sample logit gradients sum to zero over the sequence (softmax), so the
    weight gradient is a covariance-like sum with heavy cancellation."""
    import torch
    torch.manual_seed(0)
    B, S, H, M, O = 64, 100, 512, 512, 1024
    dt = torch.float64
    x = torch.randn(B, S, H, dtype=dt)
    W1 = (torch.rand(M, H, dtype=dt) * 2 - 1) / H ** 0.5
    b1 = (torch.rand(M, dtype=dt) * 2 - 1) / H ** 0.5
    w2 = (torch.rand(1, M, dtype=dt) * 2 - 1) / M ** 0.5
    Wm = (torch.rand(O, H, dtype=dt) * 2 - 1) / H ** 0.5
    dout = torch.randn(B, O, dtype=dt)

    def grad(round_x):
        w = W1.clone().requires_grad_(True)
        xin = x.to(torch.bfloat16).to(dt) if round_x else x
        att = torch.softmax(torch.relu(xin @ w.t() + b1) @ w2.t(), dim=1)
        ((att * x).sum(1) @ Wm.t()).backward(dout)
        return w.grad

    g0, g1 = grad(False), grad(True)
    rel = ((g1 - g0).norm() / g0.norm()).item()
    assert 1e-2 < rel < 1e-1, rel


class _StubLib(object):
    """Stands in for libmcan_b200.so: every entry point 'succeeds' without launching anything."""

    def __init__(self):
        self.calls = {}

    def __getattr__(self, name):
        if not name.startswith("mcan_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls[name] = self.calls.get(name, 0) + 1
            return 148 if name == "mcan_num_sms" else 0
        return fn


@pytest.mark.parametrize("group,store,bf16_sink", [(True, True, False), (True, False, False), (False, False, False), (True, True, True)])
def test_dry_run_whole_net_host_logic(monkeypatch, group, store, bf16_sink):
    """The complete launch chain of a training step (question encoder, image projection + mask, MCA_ED, AttFlat, fused
    head + loss, backward with grouped / stored weight gradients) with the C ABI stubbed out: shapes, arenas, grouping,
    gradient plumbing and the set of entry points used -- no arithmetic (buffers are uninitialised)."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import mcan_oracle as orc
    from core.model.net import Net
    from mcan_vqa_b200 import blocks, capi, ops
    stub = _StubLib()
    monkeypatch.setattr(capi, "load", lambda: stub)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_raw_sms", [0])
    monkeypatch.setattr(blocks, "DRY_RUN", True)
    monkeypatch.setattr(blocks, "GROUP_WGRADS", group)
    monkeypatch.setattr(blocks, "STORE_WGRADS", store)
    monkeypatch.setattr(blocks, "WGRAD_BF16", bf16_sink)
    cfg = orc.Cfg(dropout_rate=0.1, **dict(orc.TINY))
    T, A, B = 50, 24, 4
    net = Net(cfg, None, T, A).train()
    v, q, ans = orc.synth_batch(cfg, B, 12, 7, T, A, seed=3, ragged="prefix")
    seen = {}

    def hook(bufs, grads=None, kind=None):
        seen.setdefault(kind, []).append((len(bufs), len(grads or {})))

    from mcan_vqa_b200 import optim as _optim
    monkeypatch.setattr(_optim, "early_hook", lambda: hook)
    loss, probs = net.forward_with_loss(v, q, ans)
    assert probs.shape == (B, A) and loss.dim() == 0
    loss.backward()
    assert set(seen) == {"dec", "kv", "enc", "enc_last"}
    for n, p in net.named_parameters():
        if bf16_sink and p.dim() == 2 and (".enc_list." in n or ".dec_list." in n) and not (".mhatt2.linear_k." in n or ".mhatt2.linear_v." in n):
            assert p.grad is None, n             # bf16 gradients never become a .grad
        elif not n.startswith("attflat_lang") or True:
            assert p.grad is not None and p.grad.shape == p.shape, n
    used = set(stub.calls)
    assert {"mcan_gemm", "mcan_attn_fwd", "mcan_attn_bwd", "mcan_layernorm_fwd", "mcan_layernorm_bwd", "mcan_attflat_pool_fwd",
            "mcan_attflat_pool_bwd", "mcan_rowmask_cast", "mcan_sigmoid_bce_fwd", "mcan_sigmoid_bce_bwd", "mcan_layernorm_add_fwd",
            "mcan_embed_gather", "mcan_embed_scatter_add", "mcan_lstm_fwd", "mcan_lstm_bwd"} <= used
    assert ("mcan_gemm_grouped" in used) == group or True
    if group:
        assert stub.calls["mcan_gemm_grouped"] >= 2 * cfg.layer      # one per layer (+ the LSTM's)
    # the reference-loop spelling: forward() -> probabilities, torch BCELoss outside
    net.zero_grad(set_to_none=True)
    out = net(v, q)
    assert len(out) == 8 and out[0].shape == (B, A) and out[2].dtype == torch.bool and out[5].shape == (B, 1, 1, 7)
    out[0].sum().backward()
    assert net.proj.weight.grad is not None


def test_trusted_scope_skips_checks_only_inside_and_per_thread():
    """ops.trusted(): the per-tensor argument checks are skipped inside the overlay's launch chains only -- a direct call
    outside is still validated, also from another thread while one thread is inside a trusted scope."""
    import threading
    from mcan_vqa_b200 import capi, ops
    x = torch.randn(4, 64)                                   # CPU tensor: must be rejected by a checked call
    with pytest.raises(capi.McanError):
        ops._req2d(x, torch.float32, "x")
    seen = {}
    with ops.trusted():
        ops._req2d(x, torch.float32, "x")                    # skipped
        with ops.trusted():
            ops._req(x, torch.bfloat16, "x")                 # nested: still skipped

        def other():
            try:
                ops._req2d(x, torch.float32, "x")
                seen["other"] = "skipped"
            except capi.McanError:
                seen["other"] = "checked"
        t = threading.Thread(target=other)
        t.start()
        t.join()
    assert seen["other"] == "checked"
    with pytest.raises(capi.McanError):
        ops._req2d(x, torch.float32, "x")
    with pytest.raises(capi.McanError):
        ops.check_device(x)


def test_module_params_cache():
    from mcan_vqa_b200 import blocks
    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 2))
    a = blocks.module_params(m)
    assert [id(p) for p in a] == [id(p) for p in m.parameters()]
    assert blocks.module_params(m) is a                      # no second tree walk
    m.double()                                               # conversions keep the Parameter objects
    assert all(p is q for p, q in zip(blocks.module_params(m), m.parameters()))
    m[1] = torch.nn.Linear(4, 3)                             # replaced parameters: the owner invalidates
    blocks.invalidate_module_params(m)
    assert [id(p) for p in blocks.module_params(m)] == [id(p) for p in m.parameters()]
