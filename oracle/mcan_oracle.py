"""CPU oracle of the MCAN co-attention hot path -- TEST INFRASTRUCTURE ONLY.

A functional restatement (plain torch tensor math on the CPU, fp32 or fp64, no nn.Module)
of what Originofamonia/mcan-vqa computes on the path named in BASELINE.json: MCA_ED
(SA / SGA / MHAtt / FFN / custom LayerNorm) + AttFlat, plus the thin shell around it
(make_mask, embedding + LSTM, img_feat_linear, proj_norm, proj, sigmoid, BCE(sum)) so that
whole-model logits and loss curves can be checked.  Every function cites the reference
file:line it follows.  All arithmetic of the reference lives in PyTorch itself
(torch.nn.Linear / matmul / softmax / Tensor.std / nn.LSTM, unpinned: requirements.txt lists
only spacy and numpy); the oracle uses the same primitives on the CPU.

Parity pinning: the reference ships NO tests or golden vectors for this path (SURVEY.md
section 4).  The oracle is therefore pinned against the reference itself: oracle/make_golden.py
imports the unmodified reference modules from /root/reference in the build container, runs
them in fp64 on seeded weights/inputs and commits outputs + gradient digests under
tests/golden/; tests/test_oracle_cpu.py checks this file against those fixtures (and against
the live reference when /root/reference is present).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
The product path (core/model/*, mcan-vqa_b200/*) never does, and has no CPU fallback.

Parameters are passed as a flat dict keyed exactly like the reference state_dict
(SURVEY.md appendix A), e.g. "backbone.enc_list.0.mhatt.linear_v.weight".
"""
import math

import numpy as np
import torch

# ------------------------------------------------------------------------------------------
# configuration bag (mirrors the attributes the reference modules read from `opt`,
# cfgs/base_cfgs.py:122-141, 235-239)
# ------------------------------------------------------------------------------------------


class Cfg(object):
    def __init__(self, hidden_size=512, multi_head=8, layer=6, ff_size=None, dropout_rate=0.1,
                 flat_mlp_size=512, flat_glimpses=1, flat_out_size=512, word_embed_size=300,
                 img_feat_size=2048, use_glove=False, **extra):
        self.hidden_size = hidden_size
        self.multi_head = multi_head
        self.layer = layer
        self.ff_size = ff_size if ff_size is not None else 4 * hidden_size  # base_cfgs.py:235
        self.hidden_size_head = hidden_size // multi_head                   # base_cfgs.py:239
        self.dropout_rate = dropout_rate
        self.flat_mlp_size = flat_mlp_size
        self.flat_glimpses = flat_glimpses
        self.flat_out_size = flat_out_size
        self.word_embed_size = word_embed_size
        self.img_feat_size = img_feat_size
        self.use_glove = use_glove
        self.training = False   # True: apply nn.Dropout where the reference does (CPU-baseline timing)
        for k, v in extra.items():
            setattr(self, k, v)


SMALL = dict(hidden_size=512, multi_head=8, layer=6, flat_mlp_size=512, flat_glimpses=1,
             flat_out_size=512)                                   # cfgs/small_model.yml:1-7
LARGE = dict(hidden_size=1024, multi_head=16, layer=6, flat_mlp_size=512, flat_glimpses=1,
             flat_out_size=2048)   # cfgs/large_model.yml:1-7, heads=16 per BASELINE.json
TINY = dict(hidden_size=128, multi_head=2, layer=2, flat_mlp_size=64, flat_glimpses=2,
            flat_out_size=128, word_embed_size=24, img_feat_size=40)   # golden-fixture config


# ------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------
def _drop(x, cfg):
    """nn.Dropout(cfg.dropout_rate) in training mode, identity otherwise."""
    if cfg is not None and getattr(cfg, "training", False) and cfg.dropout_rate > 0:
        return torch.nn.functional.dropout(x, cfg.dropout_rate, True)
    return x


def linear(x, w, b):
    """nn.Linear: y = x W^T + b (weights are (out, in))."""
    return torch.matmul(x, w.t()) + b


def layer_norm(x, a2, b2, eps=1e-6):
    """core/model/net_utils.py:56-60 -- unbiased std, eps added to std."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a2 * (x - mean) / (std + eps) + b2


def layer_norm_backward(dy, x, a2, eps=1e-6):
    """Closed-form backward of layer_norm; what mcan_layernorm_bwd computes.
    Returns (dx, da2, db2)."""
    n = x.shape[-1]
    mean = x.mean(-1, keepdim=True)
    c = x - mean
    sigma = x.std(-1, keepdim=True)
    s = sigma + eps
    gh = dy * a2
    dx = (gh - gh.mean(-1, keepdim=True)) / s - c * (gh * c).sum(-1, keepdim=True) / (s * s * sigma * (n - 1))
    red = tuple(range(x.dim() - 1))
    return dx, (dy * c / s).sum(red), dy.sum(red)


def attention(value, key, query, mask, dropout_keep=None, dropout_p=0.0, cfg=None):
    """core/model/mca.py:65-78.  value/key/query: [B,h,S,d]; mask: bool [B,1,1,Sk] or None.
    dropout_keep: optional bool [B,h,Sq,Sk] keep-mask standing in for nn.Dropout (mca.py:76)."""
    d_k = query.size(-1)
    scores = torch.matmul(query, key.transpose(-2, -1)) / math.sqrt(d_k)
    if mask is not None:
        scores = scores.masked_fill(mask, -1e9)
    att_map = torch.softmax(scores, dim=-1)
    if dropout_keep is not None:
        att_map = att_map * dropout_keep.to(att_map.dtype) / (1.0 - dropout_p)
    att_map = _drop(att_map, cfg)
    return torch.matmul(att_map, value)


def attention_backward(dout, value, key, query, mask):
    """Closed-form backward of attention() without dropout; what mcan_attn_bwd computes.
    Returns (dvalue, dkey, dquery)."""
    d_k = query.size(-1)
    scale = 1.0 / math.sqrt(d_k)
    scores = torch.matmul(query, key.transpose(-2, -1)) * scale
    if mask is not None:
        scores = scores.masked_fill(mask, -1e9)
    p = torch.softmax(scores, dim=-1)
    dv = torch.matmul(p.transpose(-2, -1), dout)
    dp = torch.matmul(dout, value.transpose(-2, -1))
    ds = p * (dp - (p * dp).sum(-1, keepdim=True))
    if mask is not None:
        ds = ds.masked_fill(mask, 0.0)   # masked_fill blocks the gradient (also for all-masked rows)
    dq = torch.matmul(ds, key) * scale
    dk = torch.matmul(ds.transpose(-2, -1), query) * scale
    return dv, dk, dq


def mhatt(p, pre, v, k, q, mask, cfg):
    """core/model/mca.py:30-63.  NOTE the argument order (v, k, q)."""
    n = q.size(0)
    h, d = cfg.multi_head, cfg.hidden_size_head

    def split(t):
        return t.view(n, -1, h, d).transpose(1, 2)

    vv = split(linear(v, p[pre + "linear_v.weight"], p[pre + "linear_v.bias"]))
    kk = split(linear(k, p[pre + "linear_k.weight"], p[pre + "linear_k.bias"]))
    qq = split(linear(q, p[pre + "linear_q.weight"], p[pre + "linear_q.bias"]))
    atted = attention(vv, kk, qq, mask, cfg=cfg)
    atted = atted.transpose(1, 2).contiguous().view(n, -1, cfg.hidden_size)
    return linear(atted, p[pre + "linear_merge.weight"], p[pre + "linear_merge.bias"])


def mlp(p, pre, x, cfg=None):
    """core/model/net_utils.py:25-45: Linear -> ReLU -> (dropout) -> Linear."""
    hmid = _drop(torch.relu(linear(x, p[pre + "fc.linear.weight"], p[pre + "fc.linear.bias"])), cfg)
    return linear(hmid, p[pre + "linear.weight"], p[pre + "linear.bias"])


def ffn(p, pre, x, cfg=None):
    """core/model/mca.py:85-98."""
    return mlp(p, pre + "mlp.", x, cfg)


def sa(p, pre, x, x_mask, cfg):
    """core/model/mca.py:118-127 (post-LN residual blocks; dropout = identity in eval)."""
    x = layer_norm(x + _drop(mhatt(p, pre + "mhatt.", x, x, x, x_mask, cfg), cfg), p[pre + "norm1.a_2"], p[pre + "norm1.b_2"])
    x = layer_norm(x + _drop(ffn(p, pre + "ffn.", x, cfg), cfg), p[pre + "norm2.a_2"], p[pre + "norm2.b_2"])
    return x


def sga(p, pre, x, y, x_mask, y_mask, cfg):
    """core/model/mca.py:150-164: self-attention on x, then attention guided by y, then FFN."""
    x = layer_norm(x + _drop(mhatt(p, pre + "mhatt1.", x, x, x, x_mask, cfg), cfg), p[pre + "norm1.a_2"], p[pre + "norm1.b_2"])
    x = layer_norm(x + _drop(mhatt(p, pre + "mhatt2.", y, y, x, y_mask, cfg), cfg), p[pre + "norm2.a_2"], p[pre + "norm2.b_2"])
    x = layer_norm(x + _drop(ffn(p, pre + "ffn.", x, cfg), cfg), p[pre + "norm3.a_2"], p[pre + "norm3.b_2"])
    return x


def mca_ed(p, pre, x, y, x_mask, y_mask, cfg):
    """core/model/mca.py:178-186: all encoders on x, then all decoders on y guided by the FINAL x."""
    for i in range(cfg.layer):
        x = sa(p, "%senc_list.%d." % (pre, i), x, x_mask, cfg)
    for i in range(cfg.layer):
        y = sga(p, "%sdec_list.%d." % (pre, i), y, x, y_mask, x_mask, cfg)
    return x, y


def mca_classifier(p, pre, y, y_mask, cfg):
    """core/model/mca.py:200-207: SA-only stack."""
    for i in range(cfg.layer):
        y = sa(p, "%senc_list.%d." % (pre, i), y, y_mask, cfg)
    return y


def attflat(p, pre, x, x_mask, cfg):
    """core/model/net.py:38-55.  Returns (x_atted [B,O], att_w [B,S,G]); softmax over dim=1."""
    att_w = mlp(p, pre + "mlp.", x, cfg)
    att_w = att_w.masked_fill(x_mask.squeeze(1).squeeze(1).unsqueeze(2), -1e9)
    att_w = torch.softmax(att_w, dim=1)
    att_list = [torch.sum(att_w[:, :, i:i + 1] * x, dim=1) for i in range(cfg.flat_glimpses)]
    x_atted = torch.cat(att_list, dim=1)
    return linear(x_atted, p[pre + "linear_merge.weight"], p[pre + "linear_merge.bias"]), att_w


def make_mask(feature):
    """core/model/net.py:135-137: True where the feature row is all zero."""
    return (torch.sum(torch.abs(feature), dim=-1) == 0).unsqueeze(1).unsqueeze(2)


def lstm(p, pre, x):
    """nn.LSTM(num_layers=1, batch_first=True) as used at core/model/net.py:73-78,103-104:
    gates ordered (i, f, g, o); zero initial state; pads run through it (not packed)."""
    w_ih, w_hh = p[pre + "weight_ih_l0"], p[pre + "weight_hh_l0"]
    b_ih, b_hh = p[pre + "bias_ih_l0"], p[pre + "bias_hh_l0"]
    bsz, steps, _ = x.shape
    hid = w_hh.shape[1]
    h = x.new_zeros(bsz, hid)
    c = x.new_zeros(bsz, hid)
    outs = []
    for t in range(steps):
        gates = linear(x[:, t], w_ih, b_ih) + linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def net_forward(p, v, ques_ix, cfg):
    """core/model/net.py:96-131 (Net; Net2 :351-375 computes the same probs).
    Returns the 8-tuple of Net.forward."""
    q_mask = make_mask(ques_ix.unsqueeze(2))
    v_mask = make_mask(v)
    q = p["embedding.weight"][ques_ix]
    q = lstm(p, "lstm.", q)
    v = linear(v, p["img_feat_linear.weight"], p["img_feat_linear.bias"])
    q, v = mca_ed(p, "backbone.", q, v, q_mask, v_mask, cfg)
    lang, q_w = attflat(p, "attflat_lang.", q, q_mask, cfg)
    img, v_w = attflat(p, "attflat_img.", v, v_mask, cfg)
    a = layer_norm(lang + img, p["proj_norm.a_2"], p["proj_norm.b_2"])
    probs = torch.sigmoid(linear(a, p["proj.weight"], p["proj.bias"]))
    return probs, v, v_mask, v_w, q, q_mask, q_w, a


def classifier_forward(p, v, cfg):
    """core/model/net.py:155-184 (ClassifierNet.forward): image-only SA stack -> AttFlat -> proj_norm ->
    proj -> sigmoid.  Returns the reference's 5-tuple (probs, v, v_mask, v_w, a)."""
    v_mask = make_mask(v)
    v = linear(v, p["img_feat_linear.weight"], p["img_feat_linear.bias"])
    v = mca_classifier(p, "backbone.", v, v_mask, cfg)
    img, v_w = attflat(p, "attflat_img.", v, v_mask, cfg)
    a = layer_norm(img, p["proj_norm.a_2"], p["proj_norm.b_2"])
    probs = torch.sigmoid(linear(a, p["proj.weight"], p["proj.bias"]))
    return probs, v, v_mask, v_w, a


def bce_sum(probs, target):
    """torch.nn.BCELoss(reduction='sum') (core/exec.py:67): log clamped at -100 like torch."""
    lp = torch.clamp(torch.log(probs), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - probs), min=-100.0)
    return -(target * lp + (1.0 - target) * l1p).sum()


def warmup_rate(step, lr_base, data_size, batch_size):
    """core/model/optim.py:36-49 (WarmupOptimizer.rate)."""
    per = data_size / batch_size
    if step <= int(per * 1):
        return float(lr_base) * 0.25
    if step <= int(per * 2):
        return float(lr_base) * 0.5
    if step <= int(per * 3):
        return float(lr_base) * 0.75
    return float(lr_base)


# ------------------------------------------------------------------------------------------
# deterministic synthetic weights and inputs.  numpy RandomState streams are
# frozen across numpy versions, so fixtures can be regenerated bit-identically anywhere.
# ------------------------------------------------------------------------------------------
def param_shapes(cfg, token_size, answer_size, classifier=False):
    """Ordered (name, shape) list == reference state_dict."""
    H, F, M, G, O = cfg.hidden_size, cfg.ff_size, cfg.flat_mlp_size, cfg.flat_glimpses, cfg.flat_out_size
    E, I = cfg.word_embed_size, cfg.img_feat_size
    out = []

    def lin(name, o, i):
        out.append((name + ".weight", (o, i)))
        out.append((name + ".bias", (o,)))

    def att(pre):
        for nm in ("linear_v", "linear_k", "linear_q", "linear_merge"):
            lin(pre + nm, H, H)

    def ffn_(pre):
        lin(pre + "mlp.fc.linear", F, H)
        lin(pre + "mlp.linear", H, F)

    def norm(pre, n):
        out.append((pre + ".a_2", (n,)))
        out.append((pre + ".b_2", (n,)))

    if not classifier:
        out.append(("embedding.weight", (token_size, E)))
        out.append(("lstm.weight_ih_l0", (4 * H, E)))
        out.append(("lstm.weight_hh_l0", (4 * H, H)))
        out.append(("lstm.bias_ih_l0", (4 * H,)))
        out.append(("lstm.bias_hh_l0", (4 * H,)))
    lin("img_feat_linear", H, I)
    for i in range(cfg.layer):
        pre = "backbone.enc_list.%d." % i
        att(pre + "mhatt.")
        ffn_(pre + "ffn.")
        norm(pre + "norm1", H)
        norm(pre + "norm2", H)
    if not classifier:
        for i in range(cfg.layer):
            pre = "backbone.dec_list.%d." % i
            att(pre + "mhatt1.")
            att(pre + "mhatt2.")
            ffn_(pre + "ffn.")
            norm(pre + "norm1", H)
            norm(pre + "norm2", H)
            norm(pre + "norm3", H)
    for nm in ("attflat_img", "attflat_lang"):
        lin(nm + ".mlp.fc.linear", M, H)
        lin(nm + ".mlp.linear", G, M)
        lin(nm + ".linear_merge", O, H * G)
    norm("proj_norm", O)
    lin("proj", answer_size, O)
    return out


def synth_state_dict(cfg, token_size, answer_size, seed=0, dtype=torch.float32, classifier=False):
    """Random-init weights with nn.Linear-like scale (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)));
    LayerNorm gains ~1, biases small but non-zero so that every term is exercised."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in param_shapes(cfg, token_size, answer_size, classifier):
        if name.endswith(".a_2"):
            arr = 1.0 + 0.1 * rs.standard_normal(shape)
        elif name.endswith(".b_2"):
            arr = 0.1 * rs.standard_normal(shape)
        elif name == "embedding.weight":
            arr = rs.standard_normal(shape)
        else:
            fan_in = shape[1] if len(shape) == 2 else {"lstm.bias_ih_l0": cfg.hidden_size,
                                                      "lstm.bias_hh_l0": cfg.hidden_size}.get(name, None)
            if fan_in is None:   # a Linear bias: fan_in of its weight = previous entry's shape[1]
                fan_in = prev_fan_in
            bound = 1.0 / math.sqrt(fan_in)
            arr = rs.uniform(-bound, bound, size=shape)
        if len(shape) == 2:
            prev_fan_in = shape[1]
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)
    return sd


def synth_batch(cfg, batch, regions, tokens, token_size, answer_size, seed=1234, ragged="none",
                dtype=torch.float32):
    """Synthetic inputs of the reference's input contract (core/data/load_data.py:294-300):
    img_feat |N(0,1)| [B,P,I] (BUTD features are post-ReLU), ques_ix int64 [B,T], soft targets.
    ragged: "none" | "prefix" (n_v~U{10..P}, n_q~U{1..T}; trailing rows / tokens zero) |
            "random" (30 % random region rows zeroed, load_data.py:239-243, plus prefix tokens)."""
    rs = np.random.RandomState(seed)
    v = np.abs(rs.standard_normal((batch, regions, cfg.img_feat_size)))
    q = rs.randint(1, token_size, size=(batch, tokens)).astype(np.int64)
    if ragged in ("prefix", "random"):
        n_q = rs.randint(1, tokens + 1, size=batch)
        for b in range(batch):
            q[b, n_q[b]:] = 0
    if ragged == "prefix":
        n_v = rs.randint(min(10, regions), regions + 1, size=batch)
        for b in range(batch):
            v[b, n_v[b]:] = 0.0
    elif ragged == "random":
        drop = rs.uniform(size=(batch, regions)) < 0.3
        v[drop] = 0.0
    ans = np.zeros((batch, answer_size))
    scores = np.array([0.3, 0.6, 0.9, 1.0])       # core/data/data_utils.py:154-164
    for b in range(batch):
        for _ in range(rs.randint(1, 4)):
            ans[b, rs.randint(0, answer_size)] = scores[rs.randint(0, 4)]
    return (torch.from_numpy(v).to(dtype), torch.from_numpy(q), torch.from_numpy(ans).to(dtype))


def grad_digest(t):
    """Compact, order-sensitive summary of a tensor used in the golden fixtures."""
    f = t.detach().double().reshape(-1)
    w = torch.cos(torch.arange(f.numel(), dtype=torch.float64) * 0.37 + 0.11)
    head = f[:5].tolist() + [0.0] * max(0, 5 - f.numel())
    return np.array([f.norm().item(), f.sum().item(), (f * w).sum().item()] + head)


# ------------------------------------------------------------------------------------------
# per-module golden fixtures (tests/golden/modules_tiny.npz): parameters are regenerated from
# these seeds instead of being stored
# ------------------------------------------------------------------------------------------
MODULE_SEEDS = {"ln": 21, "mhatt_self": 22, "mhatt_guided": 22, "sa": 23, "sga": 24, "attflat": 25}


def seeded_params(names, shapes, seed):
    """The parameter values of a module fixture: uniform(-0.2, 0.2) drawn in named_parameters order."""
    r = np.random.RandomState(seed)
    return {n: torch.from_numpy(r.uniform(-0.2, 0.2, size=tuple(int(d) for d in shp if d > 0)))
            for n, shp in zip(names, shapes)}
