"""LayerNorm / attention / AttFlat-pool / helper kernels through the C ABI vs torch fp32 math."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mcan_vqa_b200 import ops
    return ops


def _ln_ref(x, a, b, eps=1e-6):
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


@pytest.mark.parametrize("rows,h", [(896, 512), (6400, 1024), (64, 2048), (37, 512), (5, 64)])
def test_layernorm_fwd_bwd(rows, h):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + h)
    x = (torch.randn(rows, h, generator=g) * 2 + 0.5).cuda().requires_grad_(True)
    a = (torch.rand(h, generator=g) + 0.5).cuda().requires_grad_(True)
    b = torch.randn(h, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(rows, h, generator=g).cuda()
    y32 = torch.empty(rows, h, device="cuda")
    ybf = torch.empty(rows, h, device="cuda", dtype=torch.bfloat16)
    ylo = torch.empty(rows, h, device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(rows, device="cuda")
    sigma = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x.detach(), a.detach(), b.detach(), 1e-6, y_f32=y32, y_bf16=ybf, y_lo=ylo, mean=mean, sigma=sigma)
    ref = _ln_ref(x.double(), a.double(), b.double())
    ref.backward(dy.double())
    torch.cuda.synchronize()
    assert (y32 - ref.float()).abs().max() < 2e-5
    assert torch.equal(ybf, y32.to(torch.bfloat16))
    assert torch.equal(ylo, (y32 - ybf.float()).to(torch.bfloat16))
    assert (sigma - x.detach().std(-1)).abs().max() < 1e-5

    dx = torch.empty(rows, h, device="cuda")
    dxbf = torch.empty(rows, h, device="cuda", dtype=torch.bfloat16)
    da, db, dbias = (torch.zeros(h, device="cuda") for _ in range(3))
    ops.layernorm_bwd(dy, x.detach(), mean, sigma, a.detach(), 1e-6, dx_f32=dx, dx_bf16=dxbf, da2=da, db2=db, dbias=dbias)
    torch.cuda.synchronize()
    scale = x.grad.abs().max().item()
    assert (dx - x.grad.float()).abs().max().item() < 2e-5 * max(scale, 1.0)
    assert torch.equal(dxbf, dx.to(torch.bfloat16))
    assert (da - a.grad.float()).abs().max() < 1e-3 * max(1.0, a.grad.abs().max().item())
    assert (db - b.grad.float()).abs().max() < 1e-3 * max(1.0, b.grad.abs().max().item())
    assert (dbias - dx.sum(0)).abs().max() < 1e-3 * max(1.0, dx.sum(0).abs().max().item())


def test_layernorm_bwd_dropout_gate_matches_gemm_epilogue_indexing():
    ops = _ops()
    rows, h, p, seed = 128, 512, 0.1, 777
    g = torch.Generator().manual_seed(3)
    x = torch.randn(rows, h, generator=g).cuda()
    dy = torch.randn(rows, h, generator=g).cuda()
    a = torch.ones(h, device="cuda"); b = torch.zeros(h, device="cuda")
    mean = torch.empty(rows, device="cuda"); sigma = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, a, b, 1e-6, y_f32=torch.empty_like(x), mean=mean, sigma=sigma)
    dx = torch.empty_like(x); dxbf = torch.empty(rows, h, device="cuda", dtype=torch.bfloat16)
    dbias = torch.zeros(h, device="cuda")
    ops.layernorm_bwd(dy, x, mean, sigma, a, 1e-6, dx_f32=dx, dx_bf16=dxbf, dropout_p=p, seed=seed, dbias=dbias)
    torch.cuda.synchronize()
    keep = ops.dropout_keep_mask(rows * h, p, seed).view(rows, h).cuda()
    ref = torch.where(keep, dx / (1 - p), torch.zeros((), device="cuda"))
    assert torch.equal(dxbf, ref.to(torch.bfloat16))
    assert (dbias - ref.sum(0)).abs().max() < 1e-3


def test_layernorm_constant_row_gives_b2():
    ops = _ops()
    x = torch.full((8, 512), 3.0, device="cuda")
    a = torch.rand(512, device="cuda"); b = torch.randn(512, device="cuda")
    y = torch.empty_like(x)
    ops.layernorm_fwd(x, a, b, 1e-6, y_f32=y)
    torch.cuda.synchronize()
    assert torch.equal(y, b.expand_as(y))


def _attn_ref(q, k, v, mask, scale, keep=None, p=0.0):
    # q [B,h,Sq,d] etc., mask bool [B,Sk]
    s = torch.matmul(q, k.transpose(-2, -1)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :], -1e9)
    pr = torch.softmax(s, dim=-1)
    if keep is not None:
        pr = pr * keep / (1 - p)
    return torch.matmul(pr, v)


ATTN_CASES = [
    # batch, heads, sq, sk, d, mask kind
    (4, 8, 14, 14, 64, "prefix"),
    (3, 8, 100, 100, 64, "random"),
    (3, 8, 100, 14, 64, "prefix"),
    (2, 8, 100, 100, 128, "random"),
    (2, 16, 60, 60, 64, "none"),
    (2, 4, 1, 1, 64, "none"),
    (2, 4, 128, 128, 64, "allmasked"),
    (2, 2, 33, 77, 128, "random"),
]


def _make_attn(batch, heads, sq, sk, d, kind, seed=0):
    g = torch.Generator().manual_seed(seed + sq * 131 + sk)
    H = heads * d
    qkv = torch.randn(batch * max(sq, sk), 3 * H, generator=g).to(torch.bfloat16).cuda()
    q = qkv[: batch * sq, :H]
    k = qkv[: batch * sk, H:2 * H]
    v = qkv[: batch * sk, 2 * H:]
    if kind == "none":
        mask = None
    elif kind == "prefix":
        lens = torch.randint(1, sk + 1, (batch,), generator=g)
        mask = (torch.arange(sk)[None, :] >= lens[:, None])
    elif kind == "random":
        mask = torch.rand(batch, sk, generator=g) < 0.3
    else:
        mask = torch.rand(batch, sk, generator=g) < 0.3
        mask[0, :] = True     # fully masked sample -> uniform softmax
    return q, k, v, (None if mask is None else mask.cuda())


def _heads(t, batch, s, heads, d):
    t = t if t.dtype == torch.float64 else t.float()
    return t.reshape(batch, s, heads, d).transpose(1, 2)


@pytest.mark.parametrize("batch,heads,sq,sk,d,kind", ATTN_CASES)
def test_attention_fwd_bwd(batch, heads, sq, sk, d, kind):
    ops = _ops()
    q, k, v, mask = _make_attn(batch, heads, sq, sk, d, kind)
    H = heads * d
    scale = 1.0 / math.sqrt(d)
    mask_u8 = None if mask is None else mask.to(torch.uint8).contiguous()
    out = torch.full((batch * sq, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attn_fwd(q, k, v, mask_u8, out, batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale)
    qh = _heads(q, batch, sq, heads, d).requires_grad_(True)
    kh = _heads(k, batch, sk, heads, d).requires_grad_(True)
    vh = _heads(v, batch, sk, heads, d).requires_grad_(True)
    ref = _attn_ref(qh, kh, vh, mask, scale)
    torch.cuda.synchronize()
    got = _heads(out, batch, sq, heads, d)
    assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item()), (got - ref).abs().max().item()

    dout = (torch.randn(batch * sq, H, device="cuda") * 0.5).to(torch.bfloat16)
    dqkv = torch.zeros(batch * max(sq, sk), 3 * H, device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = dqkv[: batch * sq, :H], dqkv[: batch * sk, H:2 * H], dqkv[: batch * sk, 2 * H:]
    dbq, dbk, dbv = (torch.full((H,), 0.25, device="cuda") for _ in range(3))    # += : start from a non-zero value
    ops.attn_bwd(q, k, v, mask_u8, dout, dq, dk, dv, batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale,
                 dbq=dbq, dbk=dbk, dbv=dbv)
    ref.backward(_heads(dout, batch, sq, heads, d))
    torch.cuda.synchronize()
    for name, gt, rf, s in (("dq", dq, qh.grad, sq), ("dk", dk, kh.grad, sk), ("dv", dv, vh.grad, sk)):
        gt = _heads(gt, batch, s, heads, d)
        err = (gt - rf).abs().max().item()
        assert err < 3e-2 * max(1.0, rf.abs().max().item()), (name, err, rf.abs().max().item())
    # the fused bias gradients = column sums of the stored (bf16) gradients
    for name, db, gt, s in (("dbq", dbq, dq, sq), ("dbk", dbk, dk, sk), ("dbv", dbv, dv, sk)):
        want = 0.25 + gt[: batch * s].float().sum(0)
        assert (db - want).abs().max().item() < 1e-3 * max(1.0, want.abs().max().item()), name


def test_attention_dropout_matches_host_hash():
    ops = _ops()
    batch, heads, sq, sk, d, p, seed = 2, 8, 100, 100, 64, 0.1, 4242
    q, k, v, mask = _make_attn(batch, heads, sq, sk, d, "random")
    H = heads * d
    scale = 1.0 / math.sqrt(d)
    mask_u8 = mask.to(torch.uint8).contiguous()
    out = torch.empty(batch * sq, H, device="cuda", dtype=torch.bfloat16)
    ops.attn_fwd(q, k, v, mask_u8, out, batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale, dropout_p=p, seed=seed)
    keep = ops.dropout_keep_mask(batch * heads * sq * sk, p, seed).view(batch, heads, sq, sk).cuda().float()
    qh = _heads(q, batch, sq, heads, d).requires_grad_(True)
    kh = _heads(k, batch, sk, heads, d).requires_grad_(True)
    vh = _heads(v, batch, sk, heads, d).requires_grad_(True)
    ref = _attn_ref(qh, kh, vh, mask, scale, keep, p)
    torch.cuda.synchronize()
    assert (_heads(out, batch, sq, heads, d) - ref).abs().max().item() < 3e-2
    dout = (torch.randn(batch * sq, H, device="cuda") * 0.5).to(torch.bfloat16)
    dq, dk, dv = (torch.zeros(batch * s, H, device="cuda", dtype=torch.bfloat16) for s in (sq, sk, sk))
    ops.attn_bwd(q, k, v, mask_u8, dout, dq, dk, dv, batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale, dropout_p=p, seed=seed)
    ref.backward(_heads(dout, batch, sq, heads, d))
    torch.cuda.synchronize()
    for gt, rf, s in ((dq, qh.grad, sq), (dk, kh.grad, sk), (dv, vh.grad, sk)):
        err = (_heads(gt, batch, s, heads, d) - rf).abs().max().item()
        assert err < 3e-2 * max(1.0, rf.abs().max().item())


TC_ATTN_CASES = [
    # batch, heads, sq, sk, mask kind, dropout: shapes the tcgen05 path takes (head dim 64, >= 49 queries, >= 33 keys)
    (3, 8, 100, 100, "none", 0.0),
    (3, 8, 100, 100, "allmasked", 0.1),     # one fully masked sample -> uniform probabilities (mca.py:70-75)
    (2, 16, 60, 60, "prefix", 0.1),
    (2, 4, 128, 128, "random", 0.1),
    (2, 4, 64, 77, "random", 0.1),          # odd key count: dropout element pairs straddle rows
    (2, 4, 99, 51, "random", 0.1),
    (2, 4, 50, 33, "prefix", 0.0),
    (64, 16, 100, 100, "none", 0.1),        # the MCAN-large launch (1024 CTAs, several waves)
]


@pytest.mark.parametrize("batch,heads,sq,sk,kind,p", TC_ATTN_CASES)
def test_attention_tcgen05_path(batch, heads, sq, sk, kind, p):
    """tcgen05 / TMEM attention (csrc/attention_tc.cu) vs fp32 torch math with the host mirror of the dropout mask, and
    vs the mma.sync kernels on the same inputs (same mask, same masking semantics)."""
    ops = _ops()
    d, seed = 64, 977
    q, k, v, mask = _make_attn(batch, heads, sq, sk, d, kind)
    H = heads * d
    scale = 1.0 / math.sqrt(d)
    mask_u8 = None if mask is None else mask.to(torch.uint8).contiguous()
    dout = (torch.randn(batch * sq, H, device="cuda") * 0.5).to(torch.bfloat16)
    kw = dict(batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale, dropout_p=p, seed=seed)
    res = {}
    try:
        for impl in (1, 0):
            ops.set_attn_impl(impl)
            out = torch.full((batch * sq, H), float("nan"), device="cuda", dtype=torch.bfloat16)
            dqkv = torch.full((batch * max(sq, sk), 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
            dq, dk, dv = dqkv[: batch * sq, :H], dqkv[: batch * sk, H:2 * H], dqkv[: batch * sk, 2 * H:]
            ops.attn_fwd(q, k, v, mask_u8, out, **kw)
            ops.attn_bwd(q, k, v, mask_u8, dout, dq, dk, dv, **kw)
            torch.cuda.synchronize()
            res[impl] = (out, dq, dk, dv)
    finally:
        ops.set_attn_impl(1)
    keep = None
    if p > 0:
        keep = ops.dropout_keep_mask(batch * heads * sq * sk, p, seed).view(batch, heads, sq, sk).cuda().float()
    qh = _heads(q, batch, sq, heads, d).requires_grad_(True)
    kh = _heads(k, batch, sk, heads, d).requires_grad_(True)
    vh = _heads(v, batch, sk, heads, d).requires_grad_(True)
    ref = _attn_ref(qh, kh, vh, mask, scale, keep, p) if p > 0 else _attn_ref(qh, kh, vh, mask, scale)
    ref.backward(_heads(dout, batch, sq, heads, d))
    refs = (ref.detach(), qh.grad, kh.grad, vh.grad)
    for i, (name, s) in enumerate((("out", sq), ("dq", sq), ("dk", sk), ("dv", sk))):
        rf = refs[i]
        tc = _heads(res[1][i], batch, s, heads, d)
        mma = _heads(res[0][i], batch, s, heads, d)
        tol = 3e-2 * max(1.0, rf.abs().max().item())
        assert not torch.isnan(tc).any(), name
        assert (tc - rf).abs().max().item() < tol, (name, (tc - rf).abs().max().item())
        assert (tc - mma).abs().max().item() < tol, (name, (tc - mma).abs().max().item())


@pytest.mark.parametrize("batch,s,h,mlp,glimpses", [(8, 100, 512, 512, 1), (4, 14, 1024, 512, 1), (3, 60, 512, 512, 2), (2, 100, 512, 256, 3),
                                                    (64, 100, 1024, 512, 1), (64, 14, 1024, 512, 1), (1, 100, 2048, 512, 2),
                                                    (5, 3, 40, 16, 5), (300, 100, 512, 512, 1), (2, 1, 8, 8, 1)])
def test_attflat_pool_fwd_bwd(batch, s, h, mlp, glimpses):
    ops = _ops()
    g = torch.Generator().manual_seed(batch * 7 + s)
    hmid = torch.relu(torch.randn(batch * s, mlp, generator=g)).to(torch.bfloat16).cuda()
    w2 = (torch.randn(glimpses, mlp, generator=g) * 0.1).cuda().requires_grad_(True)
    b2 = torch.randn(glimpses, generator=g).cuda().requires_grad_(True)
    x = torch.randn(batch * s, h, generator=g).cuda().requires_grad_(True)
    mask = torch.rand(batch, s, generator=g) < 0.3
    mask[0, :] = True if batch > 2 else mask[0, :]
    mask = mask.cuda()
    att_w = torch.empty(batch, s, glimpses, device="cuda")
    p32 = torch.empty(batch, glimpses * h, device="cuda")
    pbf = torch.empty(batch, glimpses * h, device="cuda", dtype=torch.bfloat16)
    mu8 = mask.to(torch.uint8).contiguous()
    ops.attflat_pool_fwd(hmid, w2.detach(), b2.detach(), mu8, x.detach(), batch=batch, s=s, h=h, mlp=mlp,
                         glimpses=glimpses, att_w=att_w, pooled_f32=p32, pooled_bf16=pbf)
    hm = hmid.float().view(batch, s, mlp).requires_grad_(True)
    logit = hm @ w2.t() + b2
    logit = logit.masked_fill(mask[:, :, None], -1e9)
    aw = torch.softmax(logit, dim=1)
    pooled = torch.cat([(aw[:, :, i:i + 1] * x.view(batch, s, h)).sum(1) for i in range(glimpses)], dim=1)
    torch.cuda.synchronize()
    assert (att_w - aw).abs().max() < 1e-5
    assert (p32 - pooled).abs().max() < 1e-4
    assert torch.equal(pbf, p32.to(torch.bfloat16))

    dpooled = torch.randn(batch, glimpses * h, generator=g).cuda()
    pooled.backward(dpooled)
    dx = torch.empty(batch * s, h, device="cuda")
    dh = torch.empty(batch * s, mlp, device="cuda", dtype=torch.bfloat16)
    dw2 = torch.zeros(glimpses, mlp, device="cuda"); db2 = torch.zeros(glimpses, device="cuda")
    ops.attflat_pool_bwd(dpooled, p32, hmid, w2.detach(), mu8, x.detach(), att_w, batch=batch, s=s, h=h, mlp=mlp,
                         glimpses=glimpses, gate_scale=1.0, dx=dx, dhmid=dh, dw2=dw2, db2=db2)
    torch.cuda.synchronize()
    assert (dx - x.grad).abs().max() < 1e-4 * max(1.0, x.grad.abs().max().item())
    ref_dh = torch.where(hmid.float().view(batch, s, mlp) > 0, hm.grad, torch.zeros((), device="cuda")).view(batch * s, mlp)
    assert (dh.float() - ref_dh).abs().max() < 1e-2 * max(1e-3, ref_dh.abs().max().item())
    assert (dw2 - w2.grad).abs().max() < 1e-3 * max(1.0, w2.grad.abs().max().item())
    assert (db2 - b2.grad).abs().max() < 1e-3 * max(1.0, b2.grad.abs().max().item())


def test_cast_and_colsum():
    ops = _ops()
    x = torch.randn(1000, 1027, device="cuda").contiguous()
    hi = torch.empty_like(x, dtype=torch.bfloat16); lo = torch.empty_like(x, dtype=torch.bfloat16)
    ops.cast_bf16(x, hi, lo)
    torch.cuda.synchronize()
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert torch.equal(lo, (x - hi.float()).to(torch.bfloat16))
    y = torch.randn(6400, 1536, device="cuda")
    out = torch.zeros(1536, device="cuda")
    ops.colsum(y, out)
    ybf = y.to(torch.bfloat16)
    out2 = torch.zeros(512, device="cuda")
    ops.colsum(ybf[:, 512:1024], out2)
    torch.cuda.synchronize()
    assert (out - y.sum(0)).abs().max() < 1e-2
    assert (out2 - ybf[:, 512:1024].float().sum(0)).abs().max() < 1e-2


@pytest.mark.parametrize("batch,heads,sq,sk,d", [(2, 8, 100, 100, 64), (2, 4, 100, 14, 64), (2, 2, 60, 60, 128)])
def test_attention_fwd_split_precision(batch, heads, sq, sk, d):
    """hi/lo operands: the kernel must reproduce fp32 attention to ~1e-5, far below bf16 rounding."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    H = heads * d
    q32 = torch.randn(batch * sq, H, generator=g).cuda()
    k32 = torch.randn(batch * sk, H, generator=g).cuda()
    v32 = torch.randn(batch * sk, H, generator=g).cuda()
    mask = (torch.rand(batch, sk, generator=g) < 0.3).cuda()
    parts = {}
    for nm, t in (("q", q32), ("k", k32), ("v", v32)):
        hi = torch.empty_like(t, dtype=torch.bfloat16); lo = torch.empty_like(t, dtype=torch.bfloat16)
        ops.cast_bf16(t, hi, lo)
        parts[nm] = (hi, lo)
    out = torch.empty(batch * sq, H, device="cuda", dtype=torch.bfloat16)
    out_lo = torch.empty_like(out)
    ops.attn_fwd(parts["q"][0], parts["k"][0], parts["v"][0], mask.to(torch.uint8).contiguous(), out, batch=batch,
                 heads=heads, sq=sq, sk=sk, head_dim=d, scale=1.0 / math.sqrt(d),
                 q_lo=parts["q"][1], k_lo=parts["k"][1], v_lo=parts["v"][1], out_lo=out_lo)
    ref = _attn_ref(_heads(q32.double(), batch, sq, heads, d).double(), _heads(k32.double(), batch, sk, heads, d).double(),
                    _heads(v32.double(), batch, sk, heads, d).double(), mask, 1.0 / math.sqrt(d))
    torch.cuda.synchronize()
    got = _heads(out, batch, sq, heads, d).double() + _heads(out_lo, batch, sq, heads, d).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-5, err


# ---- the two ends of the path: zero-row mask + cast, LayerNorm of a sum, sigmoid + BCE(sum) ----------------
@pytest.mark.parametrize("rows,cols", [(6400, 2048), (300, 2048), (37, 300), (5, 7), (64, 1026)])
def test_rowmask_cast(rows, cols):
    """reference net.py:135-137: mask = (sum(|x|, -1) == 0); the same pass writes the bf16 operand (hi, lo)."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows * 7 + cols)
    x = torch.randn(rows, cols, generator=g)
    zero_rows = torch.rand(rows, generator=g) < 0.3
    x[zero_rows] = 0.0
    x[0] = 0.0
    x[0, cols - 1] = 1e-30 if rows > 1 else 0.0          # a single tiny entry: NOT a padded row
    if rows > 2:
        x[2] = 0.0
        x[2, 0] = -0.0                                    # negative zero is zero
    x = x.cuda()
    ld = (cols + 7) // 8 * 8
    hi = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)[:, :cols]
    lo = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)[:, :cols]
    mask = torch.full((rows,), 7, device="cuda", dtype=torch.uint8)
    ops.rowmask_cast(x, hi, lo, mask)
    torch.cuda.synchronize()
    ref_mask = (x.abs().sum(-1) == 0)
    assert torch.equal(mask.bool(), ref_mask)
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert torch.equal(lo, (x - hi.float()).to(torch.bfloat16))


@pytest.mark.parametrize("rows,h", [(64, 1024), (64, 2048), (3, 512)])
def test_layernorm_of_a_sum(rows, h):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + h)
    x = torch.randn(rows, h, generator=g).cuda()
    x2 = (torch.randn(rows, h, generator=g) * 3).cuda()
    a = (torch.rand(h, generator=g) + 0.5).cuda()
    b = torch.randn(h, generator=g).cuda()
    s = torch.empty(rows, h, device="cuda")
    y32 = torch.empty(rows, h, device="cuda")
    ybf = torch.empty(rows, h, device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(rows, device="cuda")
    sigma = torch.empty(rows, device="cuda")
    ops.layernorm_add_fwd(x, x2, a, b, 1e-6, s_out=s, y_f32=y32, y_bf16=ybf, mean=mean, sigma=sigma)
    torch.cuda.synchronize()
    assert torch.equal(s, x + x2)
    ref = _ln_ref((x + x2).double(), a.double(), b.double())
    assert (y32 - ref.float()).abs().max() < 3e-5
    assert torch.equal(ybf, y32.to(torch.bfloat16))
    assert (sigma - (x + x2).double().std(-1).float()).abs().max() < 1e-4


@pytest.mark.parametrize("rows,cols", [(64, 3129), (8, 3129), (64, 100), (1, 5), (300, 777)])
def test_sigmoid_bce_sum_fwd_bwd(rows, cols):
    """reference net.py:129 (sigmoid) + exec.py:67,178 (BCELoss(reduction='sum')) and their autograd backward."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows * 31 + cols)
    ld = (cols + 3) // 4 * 4
    z = (torch.randn(rows, cols, generator=g) * 4)
    z[0, 0] = 60.0          # saturated: p == 1, log(1 - p) clamps at -100
    z[-1, -1] = -120.0      # p == 0 (denormal underflow), log p clamps at -100
    t = torch.rand(rows, cols, generator=g)
    t[torch.rand(rows, cols, generator=g) < 0.8] = 0.0
    zbuf = torch.zeros(rows, ld, device="cuda")
    zbuf[:, :cols] = z.cuda()
    logits = zbuf[:, :cols]
    tc = t.cuda()
    probs = torch.empty(rows, cols, device="cuda")
    loss = torch.full((), -1.0, device="cuda")
    for _ in range(2):      # twice: the ticket in the workspace must be left ready for the next call
        ops.sigmoid_bce_fwd(logits, probs, tc, loss)
    zr = z.cuda().requires_grad_(True)
    pr = torch.sigmoid(zr)
    lr = torch.nn.BCELoss(reduction="sum")(pr, tc)
    (lr * 0.37).backward()
    torch.cuda.synchronize()
    assert (probs - pr.detach()).abs().max() < 2e-7
    ref64 = torch.nn.BCELoss(reduction="sum")(torch.sigmoid(z.double()).float().double(), t.double())
    assert abs(loss.item() - lr.item()) <= 2e-6 * abs(lr.item()) + 1e-4, (loss.item(), lr.item(), ref64.item())
    ldz = (cols + 7) // 8 * 8
    dz = torch.full((rows, ldz), 3.0, device="cuda", dtype=torch.bfloat16)
    dbias = torch.zeros(cols, device="cuda")
    gs = torch.full((), 0.37, device="cuda")
    ops.sigmoid_bce_bwd(probs, dz[:, :cols], target=tc, gscale=gs, dbias=dbias)
    torch.cuda.synchronize()
    assert torch.equal(dz[:, cols:], torch.zeros_like(dz[:, cols:]))
    ref = zr.grad
    assert (dz[:, :cols].float() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6
    assert (dbias - dz[:, :cols].float().sum(0)).abs().max() < 1e-4 * max(1.0, dz.float().abs().max().item()) * rows ** 0.5
    # plain sigmoid backward of an element-wise gradient (the reference loop: torch's BCELoss on forward()'s probs)
    gout = torch.randn(rows, cols, generator=g).cuda()
    zr.grad = None
    (torch.sigmoid(zr) * gout).sum().backward()
    ops.sigmoid_bce_bwd(probs, dz[:, :cols], gout=gout)
    torch.cuda.synchronize()
    assert (dz[:, :cols].float() - zr.grad).abs().max() <= 2 ** -8 * zr.grad.abs().max() + 1e-6
    # probabilities only (no target, no loss)
    probs2 = torch.empty_like(probs)
    ops.sigmoid_bce_fwd(logits, probs2)
    torch.cuda.synchronize()
    assert torch.equal(probs, probs2)


# ---- question encoder: embedding + LSTM (csrc/lstm.cu) vs torch.nn.Embedding / nn.LSTM in fp64 ------------------
@pytest.mark.parametrize("B,T,E,H,V", [(64, 14, 300, 1024, 2000), (64, 14, 300, 512, 2000), (5, 14, 300, 128, 50),
                                       (70, 9, 300, 256, 300), (1, 1, 64, 128, 10), (130, 14, 300, 512, 500)])
def test_question_encoder_embedding_lstm_fwd_bwd(B, T, E, H, V):
    """reference net.py:66-78, 99, 103-104: gates (i, f, g, o), zero initial state, pads run through the LSTM;
    bf16 operands with fp32 accumulation and fp32 cell state (tolerances as for the other bf16 GEMM chains)."""
    from mcan_vqa_b200 import blocks
    from mcan_vqa_b200.blocks import LinearParams, Runtime
    torch.manual_seed(B * 1000 + H)
    emb = torch.nn.Embedding(V, E).cuda()
    lstm = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).cuda()
    tokens = torch.randint(1, V, (B, T), device="cuda")
    tokens[:, T - T // 3:] = 0                   # padding tokens (index 0), as in VQA questions
    tokens[0] = 0                                # a fully padded question
    assert blocks.lstm_supported(lstm, False)
    lp_ih = LinearParams([(lstm.weight_ih_l0, lstm.bias_ih_l0)], pad=64).get(True, True)      # hi + lo: split-precision input projection
    lp_hh = LinearParams([(lstm.weight_hh_l0, lstm.bias_hh_l0)]).get(True)
    rt = Runtime(True, 0.0)
    q, mask, ctx = blocks.qenc_fwd(rt, emb.weight.detach(), lp_ih, lp_hh, tokens, True)
    dq = torch.randn(B * T, H, device="cuda") * 0.1
    dt, dwi, dbi, dwh, dbh = blocks.qenc_bwd(rt, ctx, dq, V)
    torch.cuda.synchronize()
    # fp64 reference
    emb64 = torch.nn.Embedding(V, E).double()
    lstm64 = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).double()
    emb64.load_state_dict({k: v.double().cpu() for k, v in emb.state_dict().items()})
    lstm64.load_state_dict({k: v.double().cpu() for k, v in lstm.state_dict().items()})
    ref, _ = lstm64(emb64(tokens.cpu()))
    ref.backward(dq.double().cpu().view(B, T, H))
    assert torch.equal(mask.view(B, T).bool().cpu(), tokens.cpu() == 0)
    assert (q.view(B, T, H).double().cpu() - ref).abs().max().item() < 1.5e-2

    def rel(a, b):
        return ((a.double().cpu() - b).norm() / (b.norm() + 1e-30)).item()
    assert rel(dwh, lstm64.weight_hh_l0.grad) < 3e-2
    assert rel(dwi, lstm64.weight_ih_l0.grad) < 3e-2
    assert rel(dbh, lstm64.bias_hh_l0.grad) < 3e-2 and rel(dbi, lstm64.bias_ih_l0.grad) < 3e-2
    assert rel(dt, emb64.weight.grad) < 3e-2
    # inference: no saved state, same outputs
    q2, _, _ = blocks.qenc_fwd(Runtime(False, 0.0), emb.weight.detach(), lp_ih, lp_hh, tokens, False)
    torch.cuda.synchronize()
    assert torch.equal(q, q2)
