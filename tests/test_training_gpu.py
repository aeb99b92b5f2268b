"""Training-level parity: loss curves over 100 optimiser steps vs the CPU oracle, CUDA-graph replay
vs eager launches, gradient accumulation, the reference's training-loop call sequence, and
size-independent properties at the BASELINE.json sizes."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mcan_oracle as orc  # noqa: E402


def _oracle_losses(cfg, sd, batches, lr_base, data_size, batch_size, steps):
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(p.values()), lr=0.0, weight_decay=1e-4)
    out = []
    for s in range(steps):
        v, q, a = batches[s % len(batches)]
        for g in opt.param_groups:
            g["lr"] = orc.warmup_rate(s + 1, lr_base, data_size, batch_size)
        opt.zero_grad()
        loss = orc.bce_sum(orc.net_forward(p, v, q, cfg)[0], a)
        loss.backward()
        opt.step()
        out.append(loss.item())
    return out


def test_loss_curve_100_steps_matches_oracle():
    """BASELINE north star: loss curves within tolerance over 100 steps (bf16 mode: 2e-2 relative).
    data_size = 25 batches so all four warm-up learning-rate levels are exercised (optim.py:36-49)."""
    from mcan_vqa_b200.train import Trainer
    cfg = orc.Cfg(dropout_rate=0.0, **orc.TINY)
    T, A, B, steps = 50, 24, 8, 100
    sd = orc.synth_state_dict(cfg, T, A, seed=0)
    batches = [orc.synth_batch(cfg, B, 12, 7, T, A, seed=100 + i, ragged="prefix") for i in range(5)]
    lr_base, data_size = 1e-3, 25 * B
    ref = _oracle_losses(cfg, sd, batches, lr_base, data_size, B, steps)
    tr = Trainer(cfg, T, A, torch.device("cuda"), lr_base=lr_base, data_size=data_size, batch_size=B, state_dict=sd)
    dev_batches = [tuple(t.cuda() for t in b) for b in batches]
    got = [tr.step(*dev_batches[s % 5]).item() for s in range(steps)]
    tr.close()
    assert ref[-1] < 0.7 * ref[0]            # the model is actually learning
    worst = max(abs(g - r) / abs(r) for g, r in zip(got, ref))
    assert worst < 2e-2, worst


def test_cuda_graph_replay_matches_eager_and_seed_word_changes_masks():
    from mcan_vqa_b200.train import Trainer
    T, A, B = 50, 24, 8
    cfg0 = orc.Cfg(dropout_rate=0.0, **orc.TINY)
    sd = orc.synth_state_dict(cfg0, T, A, seed=1)
    batch = tuple(t.cuda() for t in orc.synth_batch(cfg0, B, 12, 7, T, A, seed=5))
    eager = Trainer(cfg0, T, A, torch.device("cuda"), lr_base=1e-3, data_size=B * 4, batch_size=B, state_dict=sd)
    le = [eager.step(*batch).item() for _ in range(8)]
    eager.close()
    graph = Trainer(cfg0, T, A, torch.device("cuda"), lr_base=1e-3, data_size=B * 4, batch_size=B, state_dict=sd, use_graph=True)
    graph.capture(*batch, warmup=3)          # 3 warm-up steps + 1 captured (not replayed) step
    lg = [graph.step(*batch).item() for _ in range(4)]
    graph.close()
    # capture consumed 3 optimiser steps eagerly; replays continue the same trajectory
    for a, b in zip(lg, le[3:7]):
        assert abs(a - b) < 5e-3 * abs(b), (lg, le)
    # with dropout the device-side seed word must give a different mask on every replay
    cfg1 = orc.Cfg(dropout_rate=0.3, **orc.TINY)
    g2 = Trainer(cfg1, T, A, torch.device("cuda"), lr_base=0.0, data_size=B * 4, batch_size=B, state_dict=sd, use_graph=True)
    g2.capture(*batch, warmup=3)
    losses = [g2.step(*batch).item() for _ in range(4)]
    g2.close()
    assert len(set(round(x, 3) for x in losses)) == 4, losses     # lr = 0: only the masks differ


def test_gradient_accumulation_and_batch_sharding_sum():
    """grad(batch of 8) == grad(first 4) + grad(last 4): the identity behind grad_accu_steps
    (core/exec.py:163-189) and behind the data-parallel all-reduce(SUM)."""
    from core.model.net import Net
    cfg = orc.Cfg(dropout_rate=0.0, **dict(orc.SMALL, layer=2))
    T, A = 200, 3129
    sd = orc.synth_state_dict(cfg, T, A, seed=2)
    v, q, a = (t.cuda() for t in orc.synth_batch(cfg, 8, 100, 14, T, A, seed=9, ragged="random"))
    net = Net(cfg, None, T, A)
    net.load_state_dict(sd)
    net = net.cuda().train()
    loss_fn = torch.nn.BCELoss(reduction="sum")
    loss_fn(net(v, q)[0], a).backward()
    full = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    loss_fn(net(v[:4], q[:4])[0], a[:4]).backward()
    loss_fn(net(v[4:], q[4:])[0], a[4:]).backward()       # accumulates into .grad
    for n, p in net.named_parameters():
        ref = full[n]
        err = (p.grad - ref).norm().item()
        # two bf16 evaluations of the same gradient: batch-size dependent kernel choices (cuDNN's TF32
        # LSTM, split-K atomics order, GEMM tiling) move bf16 roundings; observed up to 2.2e-2 on the
        # AttFlat MLP weight, the oracle tolerance for weight gradients is 6e-2
        assert err < 4e-2 * max(ref.norm().item(), 1e-3 * max(f.norm().item() for f in full.values())), n


def test_reference_training_loop_call_sequence_and_checkpoint_roundtrip(tmp_path):
    """The calls core/exec.py makes (exec.py:52-58,104,157-208,241-253,70-94) on the overlay classes."""
    from core.model.net import Net2
    from core.model.optim import adjust_lr, get_optim
    cfg = orc.Cfg(dropout_rate=0.1, **orc.TINY)
    cfg.lr_base, cfg.batch_size, cfg.grad_norm_clip = 1e-3, 4, -1
    T, A = 50, 24
    torch.manual_seed(7)
    net = Net2(cfg, None, T, A)
    net.cuda()
    net.train()
    loss_fn = torch.nn.BCELoss(reduction="sum").cuda()
    optim = get_optim(cfg, net, data_size=16)
    v, q, a = (t.cuda() for t in orc.synth_batch(cfg, 4, 12, 7, T, A, seed=3, ragged="prefix"))
    first = None
    for step in range(12):
        optim.zero_grad()
        pred, v_out, v_mask, q_out, q_mask = net(v, q)
        loss = loss_fn(pred, a)
        loss.backward()
        first = loss.item() if first is None else first
        norms = [torch.norm(p.grad).cpu().item() for _, p in net.named_parameters() if p.grad is not None]
        assert len(norms) == len(list(net.parameters())) and all(n == n for n in norms)
        optim.step()
    assert loss.item() < first
    adjust_lr(optim, 0.2)
    state = {"state_dict": net.state_dict(), "optimizer": optim.optimizer.state_dict(), "lr_base": optim.lr_base}
    path = str(tmp_path / "epoch1.pt")
    torch.save(state, path)
    net2 = Net2(cfg, None, T, A)
    net2.cuda().eval()
    net2.load_state_dict(torch.load(path)["state_dict"])
    net.eval()
    with torch.no_grad():
        assert torch.equal(net(v, q)[0], net2(v, q)[0])


@pytest.mark.parametrize("model", ["small", "large"])
def test_full_size_properties(model):
    """BASELINE.json shapes (batch 64, 100 regions, 14 tokens): properties that need no oracle run."""
    from core.model.net import Net
    cfgd = orc.SMALL if model == "small" else orc.LARGE
    cfg = orc.Cfg(dropout_rate=0.1, **cfgd)
    T, A, B = 20000, 3129, 64
    torch.manual_seed(0)
    net = Net(cfg, None, T, A).cuda().eval()
    v, q, a = (t.cuda() for t in orc.synth_batch(cfg, B, 100, 14, T, A, seed=1234, ragged="prefix"))
    with torch.no_grad():
        out = net(v, q)
        probs = out[0]
        assert probs.shape == (B, A) and torch.isfinite(probs).all() and (probs > 0).all() and (probs < 1).all()
        assert torch.equal(probs, net(v, q)[0])                  # eval mode is deterministic
        # samples are independent: a sub-batch gives the same answers
        sub = net(v[8:24], q[8:24])[0]
        assert (sub - probs[8:24]).abs().max() < 2e-3
        # AttFlat weights are a distribution over the valid positions
        v_w, q_w = out[3], out[6]
        assert (v_w.sum(1) - 1).abs().max() < 1e-4 and (q_w.sum(1) - 1).abs().max() < 1e-4
        assert (v_w[out[2].reshape(B, 100)] < 1e-6).all()
        # padded rows never reach the logits (SURVEY section 4 edge case): the hidden states of masked
        # positions are perturbed by +100 at the input of the backbone; every masked position is excluded
        # as a key in MHAtt and as a row in AttFlat, so the logits must not move by a single bit
        H = cfg.hidden_size
        g = torch.Generator(device="cuda").manual_seed(5)
        qh = torch.randn(B, 14, H, device="cuda", generator=g)
        vh = torch.randn(B, 100, H, device="cuda", generator=g)
        q_mask, v_mask = out[5], out[2]
        assert q_mask.any() and v_mask.any()          # the prefix-ragged batch has padding on both sides

        def head(qh, vh):
            qo, vo = net.backbone(qh, vh, q_mask, v_mask)
            lang, _ = net.attflat_lang(qo, q_mask)
            img, _ = net.attflat_img(vo, v_mask)
            return net.proj(net.proj_norm(lang + img))

        base = head(qh, vh)
        qh2 = qh + 100.0 * q_mask.reshape(B, 14, 1)
        vh2 = vh + 100.0 * v_mask.reshape(B, 100, 1)
        assert not torch.equal(qh2, qh) and not torch.equal(vh2, vh)
        assert torch.equal(head(qh2, vh2), base)
    # ... and they receive an exactly-zero gradient
    qh_g = qh.clone().requires_grad_(True)
    vh_g = vh.clone().requires_grad_(True)
    head(qh_g, vh_g).square().sum().backward()
    assert torch.isfinite(qh_g.grad).all() and torch.isfinite(vh_g.grad).all()
    assert qh_g.grad[q_mask.reshape(B, 14)].abs().max().item() == 0.0
    assert vh_g.grad[v_mask.reshape(B, 100)].abs().max().item() == 0.0
    assert qh_g.grad[~q_mask.reshape(B, 14)].abs().max().item() > 0.0
    # one training step at full size produces finite gradients for every parameter
    net.train()
    loss = torch.nn.BCELoss(reduction="sum")(net(v, q)[0], a)
    loss.backward()
    for n, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_deferred_weight_gradients_on_second_stream_match_inline():
    """MCA_ED's backward runs the decoder wgrads on a second stream next to the encoder backward
    (blocks.OVERLAP_WGRAD); every parameter gradient must equal the single-stream result."""
    from core.model.net import Net
    from mcan_vqa_b200 import blocks
    cfg = orc.Cfg(dropout_rate=0.1, **orc.TINY)
    T, A, B = 50, 24, 8
    sd = orc.synth_state_dict(cfg, T, A, seed=11)
    v, q, a = (t.cuda() for t in orc.synth_batch(cfg, B, 12, 7, T, A, seed=12, ragged="prefix"))
    got = {}
    saved = blocks.OVERLAP_WGRAD
    try:
        for mode in (False, True):
            blocks.OVERLAP_WGRAD = mode
            torch.manual_seed(3)
            blocks._seed_counter[0] = 0
            net = Net(cfg, None, T, A)
            net.load_state_dict(sd)
            net = net.cuda().train()
            for _ in range(3):      # repeated use of the side stream
                net.zero_grad(set_to_none=True)
                loss = torch.nn.BCELoss(reduction="sum")(net(v, q)[0], a)
                loss.backward()
            torch.cuda.synchronize()
            got[mode] = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    finally:
        blocks.OVERLAP_WGRAD = saved
    assert got[True].keys() == got[False].keys()
    for n in got[True]:
        ref = got[False][n]
        err = (got[True][n] - ref).abs().max().item()
        assert err <= 1e-5 * (ref.abs().max().item() + 1e-12) + 1e-7, (n, err)
