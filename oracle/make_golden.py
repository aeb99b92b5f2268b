"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference

The reference modules (core/model/net.py, mca.py, net_utils.py) are imported from
/root/reference, loaded with the oracle's deterministic synthetic weights
(mcan_oracle.synth_state_dict), and run in fp64 on the CPU on seeded synthetic batches
(dense, prefix-ragged, random-ragged masks).  Stored per case: the inputs, every forward
output of Net.forward, the BCE(sum) loss and a digest (norm, sum, weighted sum, first 5
entries) of the gradient of EVERY parameter -- weights themselves are regenerated from the
seed, so the fixtures stay small.  Nothing under /root/reference is copied.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MCAN_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import mcan_oracle as orc  # noqa: E402


def import_reference():
    """Imports the reference's model modules with only the reference on sys.path."""
    saved = list(sys.path)
    for k in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import core.model.net as ref_net
        import core.model.mca as ref_mca
        import core.model.net_utils as ref_utils
    finally:
        sys.path[:] = saved
    assert ref_net.__file__.startswith(REF), ref_net.__file__
    mods = (ref_net, ref_mca, ref_utils)
    for k in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
        del sys.modules[k]
    return mods


CASES = [
    # name, cfg dict, batch, regions, tokens, token_size, answer_size, ragged, seeds (weights, batch)
    ("tiny_dense", orc.TINY, 3, 10, 6, 50, 24, "none", (0, 1234)),
    ("tiny_prefix", orc.TINY, 4, 12, 7, 50, 24, "prefix", (1, 77)),
    ("tiny_random", orc.TINY, 4, 16, 5, 50, 24, "random", (2, 78)),
    ("tiny_d128", dict(orc.TINY, hidden_size=256, multi_head=2, flat_glimpses=1), 2, 9, 4, 50, 24, "random", (3, 79)),
]


def run_case(ref_net, name, cfgd, batch, regions, tokens, token_size, answer_size, ragged, seeds):
    cfg = orc.Cfg(dropout_rate=0.0, **cfgd)
    sd = orc.synth_state_dict(cfg, token_size, answer_size, seed=seeds[0], dtype=torch.float64)
    v, q, ans = orc.synth_batch(cfg, batch, regions, tokens, token_size, answer_size, seed=seeds[1],
                                ragged=ragged, dtype=torch.float64)
    if ragged == "random":
        v[0] = 0.0      # one sample with EVERY region masked: uniform-softmax edge case
    net = ref_net.Net(cfg, None, token_size, answer_size).double()
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.eval()
    probs, v_out, v_mask, v_w, q_out, q_mask, q_w, a = net(v, q)
    loss = torch.nn.BCELoss(reduction="sum")(probs, ans)
    loss.backward()
    out = {
        "img_feat": v.numpy(), "ques_ix": q.numpy(), "ans": ans.numpy(),
        "probs": probs.detach().numpy(), "v": v_out.detach().numpy(), "q": q_out.detach().numpy(),
        "v_mask": v_mask.numpy(), "q_mask": q_mask.numpy(), "v_w": v_w.detach().numpy(),
        "q_w": q_w.detach().numpy(), "a": a.detach().numpy(), "loss": np.array(loss.item()),
        "meta": np.array([batch, regions, tokens, token_size, answer_size, seeds[0], seeds[1]]),
    }
    names = []
    digests = []
    for n, p in net.named_parameters():
        names.append(n)
        digests.append(orc.grad_digest(p.grad))
    out["grad_names"] = np.array(names)
    out["grad_digests"] = np.stack(digests)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print("%-12s loss=%.6f  params=%d  top1=%s" % (name, loss.item(), len(names), probs.argmax(1).tolist()))


def fill_seeded(mod, seed):
    names = [n for n, _ in mod.named_parameters()]
    shapes = [list(p.shape) for _, p in mod.named_parameters()]
    vals = orc.seeded_params(names, shapes, seed)
    for n, p in mod.named_parameters():
        p.data.copy_(vals[n])


def module_goldens(ref_mca, ref_utils, ref_net):
    """Per-module fixtures: LayerNorm, MHAtt, SA, SGA, AttFlat forward outputs + input grads."""
    cfg = orc.Cfg(dropout_rate=0.0, **orc.TINY)
    rs = np.random.RandomState(5)
    H = cfg.hidden_size
    out = {}
    x = torch.from_numpy(rs.standard_normal((3, 9, H))).requires_grad_(True)
    y = torch.from_numpy(rs.standard_normal((3, 5, H))).requires_grad_(True)
    x_mask = torch.from_numpy(rs.uniform(size=(3, 1, 1, 9)) < 0.3)
    y_mask = torch.from_numpy(rs.uniform(size=(3, 1, 1, 5)) < 0.3)
    x_mask[1] = True     # fully masked sample
    out.update(x=x.detach().numpy(), y=y.detach().numpy(), x_mask=x_mask.numpy(), y_mask=y_mask.numpy())

    def seeded(mod, seed):
        mod = mod.double().eval()
        fill_seeded(mod, seed)
        return mod

    def record(tag, mod, fn):
        for p in mod.parameters():
            p.grad = None
        x.grad = None
        y.grad = None
        res = fn()
        res_t = res[0] if isinstance(res, tuple) else res
        g = torch.from_numpy(np.random.RandomState(11).standard_normal(tuple(res_t.shape)))
        res_t.backward(g)
        out[tag + "_out"] = res_t.detach().numpy()
        out[tag + "_gout"] = g.numpy()
        if x.grad is not None:
            out[tag + "_dx"] = x.grad.numpy().copy()
        if y.grad is not None:
            out[tag + "_dy"] = y.grad.numpy().copy()
        # parameters are regenerated from the seed (fill_seeded); store names/shapes + grad digests
        out[tag + "_pnames"] = np.array([n for n, _ in mod.named_parameters()])
        out[tag + "_pshapes"] = np.array([list(p.shape) + [0] * (2 - p.dim()) for _, p in mod.named_parameters()])
        out[tag + "_gdigests"] = np.stack([orc.grad_digest(p.grad) for _, p in mod.named_parameters()])
        if isinstance(res, tuple):
            out[tag + "_out2"] = res[1].detach().numpy()

    ln = seeded(ref_utils.LayerNorm(H), 21)
    record("ln", ln, lambda: ln(x))
    mh = seeded(ref_mca.MHAtt(cfg), 22)
    record("mhatt_self", mh, lambda: mh(x, x, x, x_mask))
    record("mhatt_guided", mh, lambda: mh(y, y, x, y_mask))
    sa = seeded(ref_mca.SA(cfg), 23)
    record("sa", sa, lambda: sa(x, x_mask))
    sga = seeded(ref_mca.SGA(cfg), 24)
    record("sga", sga, lambda: sga(x, y, x_mask, y_mask))
    af = seeded(ref_net.AttFlat(cfg), 25)
    record("attflat", af, lambda: af(x, x_mask))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "modules_tiny.npz"), **out)
    print("modules_tiny: %d arrays" % len(out))


def classifier_golden(ref_net):
    """ClassifierNet (reference net.py:138-184): the image-only SA stack; forward 5-tuple, loss, grad digests."""
    cfg = orc.Cfg(dropout_rate=0.0, **dict(orc.TINY, layer=2))
    batch, regions, answer_size, wseed, bseed = 5, 12, 24, 21, 22
    sd = orc.synth_state_dict(cfg, 50, answer_size, seed=wseed, dtype=torch.float64, classifier=True)
    v, _, ans = orc.synth_batch(cfg, batch, regions, 7, 50, answer_size, seed=bseed, ragged="random", dtype=torch.float64)
    net = ref_net.ClassifierNet(cfg, answer_size).double()
    net.load_state_dict(sd, strict=True)
    net.eval()
    probs, v_out, v_mask, v_w, a = net(v)
    loss = torch.nn.BCELoss(reduction="sum")(probs, ans)
    loss.backward()
    used = [(n, p) for n, p in net.named_parameters() if p.grad is not None]
    out = {
        "img_feat": v.numpy(), "ans": ans.numpy(), "probs": probs.detach().numpy(), "v": v_out.detach().numpy(),
        "v_mask": v_mask.numpy(), "v_w": v_w.detach().numpy(), "a": a.detach().numpy(), "loss": np.array(loss.item()),
        "meta": np.array([batch, regions, answer_size, wseed, bseed]),
        "grad_names": np.array([n for n, _ in used]),
        "grad_digests": np.stack([orc.grad_digest(p.grad) for _, p in used]),
        "no_grad_names": np.array([n for n, p in net.named_parameters() if p.grad is None]),
    }
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "classifier_tiny.npz"), **out)
    print("classifier_tiny loss=%.6f  params with grad=%d, without=%d" % (loss.item(), len(used), len(out["no_grad_names"])))


def main():
    torch.manual_seed(0)
    ref_net, ref_mca, ref_utils = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else None       # e.g. `make_golden.py classifier`: just that fixture
    if only in (None, "net"):
        for case in CASES:
            run_case(ref_net, *case)
    if only in (None, "modules"):
        module_goldens(ref_mca, ref_utils, ref_net)
    if only in (None, "classifier"):
        classifier_golden(ref_net)


if __name__ == "__main__":
    main()
