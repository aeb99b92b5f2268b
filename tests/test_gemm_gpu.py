"""tcgen05 GEMM (mcan_gemm through the C ABI) vs torch fp32 matmul on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mcan_vqa_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(got, ref, tol, what):
    err = (got.float() - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-20
    assert err / den < tol, "%s: max abs err %.4g / max ref %.4g = %.3g" % (what, err, den, err / den)


SHAPES_NT = [
    (128, 128, 64), (128, 256, 128), (256, 128, 192), (896, 512, 512), (6400, 1024, 1024),
    (100, 512, 512),     # ragged M (TMA zero fill + row guards)
    (42, 136, 64),       # ragged M and N (scalar tail path)
    (64, 2048, 1024),    # AttFlat linear_merge shape
    (300, 1536, 2048),
]


@pytest.mark.parametrize("m,n,k", SHAPES_NT)
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_nt_plain(m, n, k, block_n, cta_group):
    ops = _ops()
    a, w = _rand((m, k), 1), _rand((n, k), 2, 0.05)
    out = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm(a, w, out_f32=out, block_n=block_n, cta_group=cta_group)
    torch.cuda.synchronize()
    _close(out, a.float() @ w.float().t(), 2e-5, "NT plain")


@pytest.mark.parametrize("m,n,k", [(256, 512, 512), (6400, 512, 2048), (100, 1024, 512)])
@pytest.mark.parametrize("cta_group,block_n", [(1, 0), (2, 128), (2, 256)])
def test_gemm_dgrad_layout(m, n, k, cta_group, block_n):
    """dX[M,K'] = dY[M,N'] W[N',K']: B is read MN-major straight from the (out,in) weight."""
    ops = _ops()
    dy, w = _rand((m, k), 3), _rand((k, n), 4, 0.05)   # contraction k = N', output n = K'
    out = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm(dy, w, b_layout=1, out_f32=out, cta_group=cta_group, block_n=block_n)
    torch.cuda.synchronize()
    _close(out, dy.float() @ w.float(), 2e-5, "dgrad layout")


@pytest.mark.parametrize("rows,n,k", [(256, 512, 512), (6400, 1024, 512), (896, 512, 2048), (100, 256, 128), (1400, 512, 512)])
@pytest.mark.parametrize("split_k", [0, 1, 3])
@pytest.mark.parametrize("cta_group,block_n", [(1, 0), (2, 128), (2, 256)])
def test_gemm_wgrad_layout_splitk(rows, n, k, split_k, cta_group, block_n):
    """dW[N',K'] = dY^T X: both operands MN-major, split-K with fp32 atomics into a zeroed buffer."""
    ops = _ops()
    dy, x = _rand((rows, n), 5, 0.1), _rand((rows, k), 6)
    out = torch.zeros((n, k), device="cuda")
    ops.gemm(dy, x, a_layout=1, b_layout=1, out_f32=out, accumulate=True, split_k=split_k, cta_group=cta_group,
             block_n=block_n)
    torch.cuda.synchronize()
    _close(out, dy.float().t() @ x.float(), 5e-5, "wgrad")


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_a_mn_b_k(cta_group):
    ops = _ops()
    at, w = _rand((320, 256), 7), _rand((384, 320), 8, 0.05)   # A stored [K,M], B stored [N,K]
    out = torch.empty((256, 384), device="cuda")
    ops.gemm(at, w, a_layout=1, b_layout=0, out_f32=out, cta_group=cta_group)
    torch.cuda.synchronize()
    _close(out, at.float().t() @ w.float().t(), 2e-5, "A MN-major, B K-major")


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_epilogue_bias_relu_resid_outputs(cta_group):
    ops = _ops()
    m, n, k = 384, 512, 256
    a, w = _rand((m, k), 9), _rand((n, k), 10, 0.1)
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda")
    o32 = torch.empty(m, n, device="cuda")
    obf = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    olo = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, relu=True, resid=resid, out_f32=o32, out_bf16=obf, out_lo=olo, cta_group=cta_group)
    torch.cuda.synchronize()
    ref = torch.relu(a.float() @ w.float().t() + bias) + resid
    _close(o32, ref, 2e-5, "bias+relu+resid fp32")
    assert torch.equal(obf, o32.to(torch.bfloat16))
    assert torch.equal(olo, (o32 - obf.float()).to(torch.bfloat16))
    _close(obf.float() + olo.float(), ref, 1e-4, "hi+lo")


def test_gemm_epilogue_gate():
    ops = _ops()
    m, n, k = 256, 384, 128
    a, w = _rand((m, k), 11), _rand((k, n), 12, 0.1)
    gate = _rand((m, n), 13)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b_layout=1, gate=gate, gate_scale=1.25, out_bf16=out)
    torch.cuda.synchronize()
    ref = torch.where(gate.float() > 0, (a.float() @ w.float()) * 1.25, torch.zeros((), device="cuda"))
    _close(out, ref, 6e-3, "gate")
    assert (out.float()[gate.float() <= 0] == 0).all()


def test_gemm_epilogue_dropout_matches_host_hash():
    ops = _ops()
    m, n, k = 256, 512, 64
    a, w = _rand((m, k), 14), _rand((n, k), 15, 0.1)
    p, seed = 0.1, 12345
    out = torch.empty(m, n, device="cuda")
    ops.gemm(a, w, dropout_p=p, seed=seed, out_f32=out)
    torch.cuda.synchronize()
    keep = ops.dropout_keep_mask(m * n, p, seed).view(m, n).cuda()
    ref = torch.where(keep, (a.float() @ w.float().t()) / (1 - p), torch.zeros((), device="cuda"))
    _close(out, ref, 2e-5, "dropout")
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p) < 0.01, frac


def test_gemm_split_precision_segments():
    """bf16x3: (A_hi,B_hi)+(A_hi,B_lo)+(A_lo,B_hi) reproduces the fp32 product to ~1e-5."""
    ops = _ops()
    m, n, k = 256, 256, 512
    g = torch.Generator().manual_seed(16)
    a32 = torch.randn(m, k, generator=g).cuda()
    w32 = (torch.randn(n, k, generator=g) * 0.05).cuda()
    ah, al = torch.empty_like(a32, dtype=torch.bfloat16), torch.empty_like(a32, dtype=torch.bfloat16)
    wh, wl = torch.empty_like(w32, dtype=torch.bfloat16), torch.empty_like(w32, dtype=torch.bfloat16)
    ops.cast_bf16(a32, ah, al)
    ops.cast_bf16(w32, wh, wl)
    out = torch.empty(m, n, device="cuda")
    ops.gemm([ah, ah, al], [wh, wl, wh], out_f32=out)
    torch.cuda.synchronize()
    ref = (a32.double() @ w32.double().t()).float()
    _close(out, ref, 3e-5, "bf16x3")
    out1 = torch.empty(m, n, device="cuda")
    ops.gemm(ah, wh, out_f32=out1)
    torch.cuda.synchronize()
    assert (out1 - ref).abs().max() > 20 * (out - ref).abs().max()


def test_gemm_strided_views_qkv():
    """Operands / outputs as column slices of wider buffers (fused QKV, cross-layer K/V)."""
    ops = _ops()
    m, h = 896, 512
    x = _rand((m, h), 17)
    wqkv = _rand((3 * h, h), 18, 0.05)
    out = torch.zeros(m, 3 * h, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, wqkv[h:2 * h], out_bf16=out[:, h:2 * h])
    torch.cuda.synchronize()
    ref = x.float() @ wqkv[h:2 * h].float().t()
    _close(out[:, h:2 * h], ref, 6e-3, "strided out")
    assert (out[:, :h] == 0).all() and (out[:, 2 * h:] == 0).all()


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (6400, 1024, 1024), (6400, 4096, 1024), (100, 512, 512), (42, 136, 64)])
@pytest.mark.parametrize("cta_group,block_n", [(1, 128), (1, 256), (2, 128), (2, 256), (0, 0)])
def test_gemm_dynamic_schedule_matches_static(m, n, k, cta_group, block_n):
    """The dynamic tile schedule (work-unit counter, used under data parallelism) computes the same
    tiles as the static one: results are bit-identical, also over repeated launches (the counter
    must reset itself)."""
    ops = _ops()
    a, w = _rand((m, k), 21), _rand((n, k), 22, 0.05)
    ref = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm(a, w, out_f32=ref, block_n=block_n, cta_group=cta_group)
    try:
        ops.set_gemm_schedule(True)
        for _ in range(3):
            out = torch.full((m, n), float("nan"), device="cuda")
            ops.gemm(a, w, out_f32=out, block_n=block_n, cta_group=cta_group)
            torch.cuda.synchronize()
            assert torch.equal(out, ref)
        acc = torch.zeros((n, k), device="cuda")
        ops.gemm(_rand((m, n), 23, 0.1), a, a_layout=1, b_layout=1, out_f32=acc, accumulate=True, split_k=3,
                 block_n=block_n, cta_group=cta_group)
        torch.cuda.synchronize()
        _close(acc, _rand((m, n), 23, 0.1).float().t() @ a.float(), 5e-5, "dynamic wgrad split-K")
    finally:
        ops.set_gemm_schedule(False)


@pytest.mark.parametrize("m,n,k", [(512, 256, 64), (6400, 1024, 1024), (896, 1024, 512), (700, 520, 192), (6400, 4096, 1024)])
@pytest.mark.parametrize("dynamic", [False, True])
def test_gemm_cluster_of_two_pairs_multicast(m, n, k, dynamic):
    """cta_group=4: clusters of two CTA pairs on a 512 x 256 super-tile, the B tile fetched once per
    cluster and TMA-multicast to both pairs.  Must equal the plain CTA-pair kernel bit for bit, in
    all three operand layouts (forward, dgrad, wgrad with split-K)."""
    ops = _ops()
    try:
        ops.set_gemm_schedule(dynamic)
        a, w = _rand((m, k), 31), _rand((n, k), 32, 0.05)
        ref = torch.full((m, n), float("nan"), device="cuda")
        out = torch.full((m, n), float("nan"), device="cuda")
        ops.gemm(a, w, out_f32=ref, block_n=256, cta_group=2)
        for _ in range(2):
            ops.gemm(a, w, out_f32=out, block_n=256, cta_group=4)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
        _close(out, a.float() @ w.float().t(), 2e-5, "cluster-4 NT")
        # dgrad layout: B MN-major
        wt = _rand((k, n), 33, 0.05)
        ops.gemm(a, wt, b_layout=1, out_f32=ref, block_n=256, cta_group=2)
        ops.gemm(a, wt, b_layout=1, out_f32=out, block_n=256, cta_group=4)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
        # wgrad layout: both MN-major, split-K atomics
        dy = _rand((m, n), 34, 0.1)
        acc = torch.zeros((n, k), device="cuda")
        ops.gemm(dy, a, a_layout=1, b_layout=1, out_f32=acc, accumulate=True, split_k=2, block_n=256, cta_group=4)
        torch.cuda.synchronize()
        _close(acc, dy.float().t() @ a.float(), 5e-5, "cluster-4 wgrad")
    finally:
        ops.set_gemm_schedule(False)


@pytest.mark.parametrize("m,n,k", SHAPES_NT + [(64, 3129, 2048), (896, 1024, 4096)])
def test_gemm_block_n_64(m, n, k):
    """128 x 64 single-CTA tiles (picked for the short-M question-side GEMMs), all three layouts,
    and the fused epilogue on an odd N (answer head: 3129 columns)."""
    ops = _ops()
    a, w = _rand((m, k), 41), _rand((n, k), 42, 0.05)
    ldo = (n + 3) // 4 * 4
    out = torch.full((m, ldo), float("nan"), device="cuda")[:, :n]
    ops.gemm(a, w, out_f32=out, block_n=64, cta_group=1)
    torch.cuda.synchronize()
    _close(out, a.float() @ w.float().t(), 2e-5, "NT 128x64")
    bias = torch.randn(n, device="cuda")
    outp = torch.full((m, ldo), float("nan"), device="cuda")[:, :n]
    ops.gemm(a, w, bias=bias, relu=True, out_f32=outp)         # auto-picked tile
    torch.cuda.synchronize()
    _close(outp, torch.relu(a.float() @ w.float().t() + bias), 2e-5, "bias+relu, auto tile")
    if n % 8 == 0:
        wt = _rand((k, n), 43, 0.05)
        ops.gemm(a, wt, b_layout=1, out_f32=out, block_n=64, cta_group=1)
        torch.cuda.synchronize()
        _close(out, a.float() @ wt.float(), 2e-5, "dgrad 128x64")
    if n % 8 == 0 and k % 8 == 0:
        dy = _rand((m, n), 44, 0.1)
        acc = torch.zeros((n, k), device="cuda")
        ops.gemm(dy, a, a_layout=1, b_layout=1, out_f32=acc, accumulate=True, block_n=64, cta_group=1)
        torch.cuda.synchronize()
        _close(acc, dy.float().t() @ a.float(), 5e-5, "wgrad 128x64")


@pytest.mark.parametrize("m,n,k", [(6400, 1024, 1024), (6400, 3072, 512), (6400, 4096, 256), (4800, 1024, 192)])
def test_gemm_tail_split_half_width_tiles(m, n, k):
    """With more 256-wide tiles than CTA pairs, the last partial wave runs as half-width (128) work
    units.  Results must equal the single-CTA 128-wide tiling bit for bit -- plain, fused epilogue
    (bias + ReLU + dropout + bf16 store), dgrad layout with residual, wgrad layout."""
    ops = _ops()
    a, w = _rand((m, k), 51), _rand((n, k), 52, 0.05)
    bias = torch.randn(n, device="cuda")
    ref, out = torch.empty((m, n), device="cuda"), torch.full((m, n), float("nan"), device="cuda")
    ops.gemm(a, w, out_f32=ref, block_n=128, cta_group=1)
    ops.gemm(a, w, out_f32=out, block_n=256, cta_group=2)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    rb, ob = torch.empty((m, n), device="cuda", dtype=torch.bfloat16), torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, relu=True, dropout_p=0.1, seed=5, out_bf16=rb, block_n=128, cta_group=1)
    ops.gemm(a, w, bias=bias, relu=True, dropout_p=0.1, seed=5, out_bf16=ob, block_n=256, cta_group=2)
    torch.cuda.synchronize()
    assert torch.equal(ob, rb)
    wt = _rand((k, n), 53, 0.05)
    resid = torch.randn(m, n, device="cuda")
    ops.gemm(a, wt, b_layout=1, resid=resid, out_f32=ref, block_n=128, cta_group=1)
    ops.gemm(a, wt, b_layout=1, resid=resid, out_f32=out, block_n=256, cta_group=2)
    ops.gemm(a, wt, b_layout=1, resid=resid, out_f32=ref.clone(), block_n=256, cta_group=1)   # 1-CTA 256-wide tiles split too
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    # wgrad layout without K split: D[n, k2] = dY^T X with dY [rows, n], X [rows, k2]
    rows, k2 = 512, 6400
    dy, x = _rand((rows, 6400), 54, 0.1), _rand((rows, n), 55)
    acc, accr = torch.zeros((6400, n), device="cuda"), torch.zeros((6400, n), device="cuda")
    ops.gemm(dy, x, a_layout=1, b_layout=1, out_f32=accr, accumulate=True, split_k=1, block_n=128, cta_group=1)
    ops.gemm(dy, x, a_layout=1, b_layout=1, out_f32=acc, accumulate=True, split_k=1, block_n=256, cta_group=2)
    torch.cuda.synchronize()
    assert torch.equal(acc, accr)


@pytest.mark.parametrize("m,n,k", [(896, 1024, 4096), (896, 1024, 12288), (200, 512, 2048)])
@pytest.mark.parametrize("split_k", [0, 1, 4])
def test_gemm_split_k_with_linear_fused_epilogue(m, n, k, split_k):
    """Split-K with bias + dropout + residual: out = resid + keep/(1-p) * (A W^T + bias), every split
    scales its partial sum, split 0 adds bias and residual, all via red.add into a zeroed buffer.
    Same for the dgrad layout with the ReLU/dropout gate."""
    ops = _ops()
    a, w = _rand((m, k), 61), _rand((n, k), 62, 0.05)
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda")
    ref = torch.empty((m, n), device="cuda")
    ops.gemm(a, w, bias=bias, dropout_p=0.1, seed=9, resid=resid, out_f32=ref)
    out = torch.zeros((m, n), device="cuda")
    ops.gemm(a, w, bias=bias, dropout_p=0.1, seed=9, resid=resid, out_f32=out, accumulate=True, split_k=split_k)
    torch.cuda.synchronize()
    _close(out, ref, 2e-5, "split-K fused epilogue")
    # identical dropout mask: an element equals the residual exactly where it was dropped.  A KEPT element whose value is
    # below the rounding granularity of the residual may be absorbed in one path and not in the other (the split-K path adds
    # its partial sums one by one, in an order that varies from run to run) -- such positions must be negligible in both.
    mism = (out == resid) != (ref == resid)
    if mism.any():
        assert mism.sum().item() <= 4 and (ref - resid).abs()[mism].max().item() < 1e-5 and (out - resid).abs()[mism].max().item() < 1e-5
    wt = _rand((k, n), 63, 0.05)
    gate = _rand((m, n), 64)
    ops.gemm(a, wt, b_layout=1, gate=gate, gate_scale=1.25, resid=resid, out_f32=ref)
    out.zero_()
    ops.gemm(a, wt, b_layout=1, gate=gate, gate_scale=1.25, resid=resid, out_f32=out, accumulate=True, split_k=split_k)
    torch.cuda.synchronize()
    _close(out, ref, 2e-5, "split-K gated dgrad")


@pytest.mark.parametrize("m,n,k,cta_group", [(6400, 1024, 512, 2), (896, 512, 256, 1), (100, 192, 128, 1), (300, 136, 64, 1)])
def test_gemm_epilogue_colsum_of_gated_output(m, n, k, cta_group):
    """colsum[n] += sum_m epilogue(acc)[m, n] (the bias gradient of the previous layer, fused into the dgrad
    GEMM that produces that layer's input gradient): fp32 values before the bf16 rounding, ragged M and N."""
    ops = _ops()
    a, w = _rand((m, k), 31), _rand((k, n), 32, 0.1)
    gate = _rand((m, n), 33)
    out = torch.empty(m, (n + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)[:, :n]
    cs = torch.full((n,), 0.5, device="cuda")
    ops.gemm(a, w, b_layout=1, gate=gate, gate_scale=1.25, out_bf16=out, colsum=cs, cta_group=cta_group)
    torch.cuda.synchronize()
    ref = torch.where(gate.float() > 0, (a.float() @ w.float()) * 1.25, torch.zeros((), device="cuda"))
    _close(out, ref, 6e-3, "gate")
    _close(cs, ref.sum(0) + 0.5, 2e-5 * (m ** 0.5), "colsum")


@pytest.mark.parametrize("m,n,k", [(256, 512, 512), (6400, 512, 2048), (100, 512, 192), (256, 1024, 1024), (6400, 1024, 1024),
                                   (896, 1024, 4096), (300, 1024, 512), (6400, 1024, 4096)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_ln_fused_epilogue(m, n, k, p_drop):
    """mcan_gemm_ln: LayerNorm(resid + dropout(a W^T + b)) in one kernel (cluster owns whole rows; N = 1024 exchanges
    the row statistics through DSMEM) vs torch fp64 math with the identical dropout mask, and vs the unfused chain."""
    ops = _ops()
    a, w = _rand((m, k), 41), _rand((n, k), 42, 0.05)
    g = torch.Generator(device="cpu").manual_seed(43)
    bias = torch.randn(n, generator=g).cuda()
    resid = (torch.randn(m, n, generator=g) * 1.5 + 0.3).cuda()
    a2 = (1.0 + 0.1 * torch.randn(n, generator=g)).cuda()
    b2 = (0.1 * torch.randn(n, generator=g)).cuda()
    eps, seed = 1e-6, 777
    s = torch.full((m, n), float("nan"), device="cuda")
    y32 = torch.full((m, n), float("nan"), device="cuda")
    ybf = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(m, device="cuda")
    sigma = torch.empty(m, device="cuda")
    ops.gemm_ln(a, w, bias=bias, resid=resid, ln_a2=a2, ln_b2=b2, eps=eps, dropout_p=p_drop, seed=seed, s_f32=s, y_f32=y32,
                y_bf16=ybf, mean=mean, sigma=sigma)
    torch.cuda.synchronize()
    acc = a.double() @ w.double().t() + bias.double()
    if p_drop > 0:
        keep = ops.dropout_keep_mask(m * n, p_drop, seed).view(m, n).cuda()
        acc = torch.where(keep, acc / (1 - p_drop), torch.zeros((), device="cuda", dtype=torch.float64))
    s_ref = acc + resid.double()
    mu = s_ref.mean(-1, keepdim=True)
    sd = s_ref.std(-1, keepdim=True)            # unbiased, like the reference
    y_ref = a2.double() * (s_ref - mu) / (sd + eps) + b2.double()
    _close(s, s_ref.float(), 2e-5, "s")
    _close(mean, mu.squeeze(-1).float(), 2e-5, "mean")
    _close(sigma, sd.squeeze(-1).float(), 2e-5, "sigma")
    _close(y32, y_ref.float(), 3e-5, "y fp32")
    assert torch.equal(ybf, y32.to(torch.bfloat16))
    # the unfused chain produces the same numbers (same dropout hash)
    s2 = torch.empty((m, n), device="cuda")
    ops.gemm(a, w, bias=bias, dropout_p=p_drop, seed=seed, resid=resid, out_f32=s2)
    y2 = torch.empty((m, n), device="cuda")
    ops.layernorm_fwd(s2, a2, b2, eps, y_f32=y2)
    torch.cuda.synchronize()
    _close(s, s2, 1e-6, "s vs unfused")
    _close(y32, y2, 2e-5, "y vs unfused")


def test_gemm_ln_large_mean_rows_are_stable():
    """Rows whose mean dwarfs their spread: the per-fragment exact statistics merged with Chan's formula must not
    lose the variance (a one-pass sum / sum-of-squares would)."""
    ops = _ops()
    m, n, k = 256, 1024, 64
    a = torch.zeros((m, k), device="cuda", dtype=torch.bfloat16)
    w = torch.zeros((n, k), device="cuda", dtype=torch.bfloat16)
    g = torch.Generator(device="cpu").manual_seed(5)
    resid = (1000.0 + 0.01 * torch.randn(m, n, generator=g)).cuda()
    y = torch.empty((m, n), device="cuda")
    ops.gemm_ln(a, w, bias=torch.zeros(n, device="cuda"), resid=resid, ln_a2=torch.ones(n, device="cuda"),
                ln_b2=torch.zeros(n, device="cuda"), eps=1e-6, y_f32=y)
    torch.cuda.synchronize()
    r = resid.double()
    ref = (r - r.mean(-1, keepdim=True)) / (r.std(-1, keepdim=True) + 1e-6)
    assert (y.double() - ref).abs().max().item() < 2e-2      # fp32 input quantisation of 1000 + 0.01 x is 6e-5 / 0.01


@pytest.mark.parametrize("rows,shapes,split_k", [
    (6400, [(1024, 1024), (3072, 1024), (1024, 1024), (1024, 1024), (4096, 1024), (1024, 4096)], 0),   # one decoder layer (large)
    (896, [(1024, 4096), (4096, 1024), (3072, 1024), (1024, 1024)], 0),                               # one encoder layer (large)
    (6400, [(512, 512), (1536, 512), (2048, 512), (512, 2048)], 0),                                   # small model
    (896, [(512, 512), (1536, 512)], 3),
    (1000, [(100, 300), (264, 72), (520, 1032)], 0),                                                  # ragged tiles
    (72, [(3136, 2048), (256, 256)], 1),
    (896, [(256, 256)] * 8, 0),                                                                        # the maximum group count
])
def test_gemm_grouped_wgrads(rows, shapes, split_k):
    """mcan_gemm_grouped == one mcan_gemm(a_layout=1, b_layout=1, accumulate) per group: dW_g += dY_g^T X_g; strided
    operand views (column slices of wider activation buffers) included."""
    ops = _ops()
    probs, refs = [], []
    for i, (n, k) in enumerate(shapes):
        wide_a = _rand((rows, (n + 64 + 7) // 8 * 8), 100 + i, 0.1)      # leading dimensions: multiples of 8 (TMA)
        wide_b = _rand((rows, (k + 8 + 7) // 8 * 8), 200 + i)
        dy, x = wide_a[:, 32:32 + n], wide_b[:, :k]
        out = torch.full((n, k), 0.5, device="cuda")              # += : starts from a non-zero value
        probs.append((dy, x, out))
        refs.append(0.5 + dy.float().t() @ x.float())
    ops.gemm_grouped(probs, split_k=split_k)
    torch.cuda.synchronize()
    for (dy, x, out), ref in zip(probs, refs):
        _close(out, ref, 5e-5, "grouped wgrad %s" % (tuple(out.shape),))


def test_layer_backward_with_grouped_wgrads_matches_individual_launches():
    """blocks.GROUP_WGRADS: the same gradients whether a layer's wgrads run as one grouped launch or one launch each."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import mcan_oracle as orc
    from core.model.mca import MCA_ED
    from mcan_vqa_b200 import blocks, capi
    cfg = orc.Cfg(dropout_rate=0.0, **dict(orc.SMALL, layer=2))
    torch.manual_seed(5)
    m = MCA_ED(cfg).cuda().train()
    x = torch.randn(8, 14, 512, device="cuda")
    y = torch.randn(8, 100, 512, device="cuda")
    xm = torch.zeros(8, 1, 1, 14, dtype=torch.bool, device="cuda"); xm[:, :, :, 11:] = True
    ym = (torch.rand(8, 1, 1, 100, device="cuda") < 0.2)
    res = {}
    saved = blocks.GROUP_WGRADS
    try:
        for grouped in (False, True):
            blocks.GROUP_WGRADS = grouped
            m.zero_grad(set_to_none=True)
            c0 = capi.launch_count
            xo, yo = m(x, y, xm, ym)
            (xo.sum() + (yo * yo).sum()).backward()
            torch.cuda.synchronize()
            res[grouped] = ({n: p.grad.clone() for n, p in m.named_parameters()}, capi.launch_count - c0)
    finally:
        blocks.GROUP_WGRADS = saved
    assert res[True][1] < res[False][1] - 8          # 2 x (4 - 1) + 2 x (6 - 1) launches fewer
    # (not bit-equal: the 896-row split-K GEMMs of each forward / backward accumulate with fp32 atomics, and the two
    #  runs are two such evaluations; a wrong or missing wgrad would be off by O(1))
    total = max(g.norm().item() for g in res[False][0].values())
    for n, g in res[False][0].items():
        err = (res[True][0][n] - g).norm().item()
        assert err <= 2e-2 * g.norm().item() + 1e-5 * total, (n, err, g.norm().item())


def test_gemm_grouped_bf16_outputs():
    """Grouped wgrad launch writing bf16 (the data-parallel bf16 exchange buffers): == bf16(fp32 result)."""
    ops = _ops()
    rows = 6400
    probs, refs = [], []
    for i, (n, k) in enumerate([(1024, 1024), (512, 2048), (264, 520)]):
        dy, x = _rand((rows, n), 300 + i, 0.1), _rand((rows, k), 400 + i)
        out = torch.full((n, k), float("nan"), device="cuda", dtype=torch.bfloat16)
        probs.append((dy, x, out))
        refs.append(dy.float().t() @ x.float())
    ops.gemm_grouped(probs, accumulate=False)
    torch.cuda.synchronize()
    for (dy, x, out), ref in zip(probs, refs):
        assert (out.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
