"""Per-kernel Python entry points: torch tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, the current stream); every arithmetic step is a
kernel of libmcan_b200.so.  All functions enqueue on torch's current CUDA stream and never
synchronise.  Tensors must be CUDA tensors; 2-D operands may be row-strided views
(stride(1) == 1).  Nothing here falls back to PyTorch math.
"""
import ctypes
import threading

import torch

from . import capi

_BF16 = torch.bfloat16
_F32 = torch.float32


# The eager route (the unchanged core/exec.py) pays these helpers ~1500 times per training step: torch.cuda.current_stream()
# and torch.cuda.current_device() go through several layers of Python (lazy-init checks, device-index parsing, a Stream
# object) -- the two C entry points behind them are used directly (tools/host_profile_gpu.py: 14.6 -> see DESIGN.md §8).
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _current_device():
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream():
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# Argument validation.  Calls made from inside the overlay modules' launch chains (blocks.py via autograd._Chain) are
# "trusted": every tensor there was allocated by the chain itself with the right dtype on the device of the module's
# input, which check_device() validated at the chain's entry -- the ~900 per-tensor checks of a training step are skipped
# (1.8 ms of host time on the eager route).  Direct calls of the functions below (tests, other users of the C ABI
# wrappers) are always checked; the C ABI validates pointers, alignments and shapes in every case.
_tls = threading.local()


class trusted(object):
    def __enter__(self):
        _tls.depth = getattr(_tls, "depth", 0) + 1

    def __exit__(self, *exc):
        _tls.depth -= 1


def check_device(t, name="input"):
    """Entry check of a launch chain: CUDA tensor on the current device (kernels are enqueued on the current device's
    current stream)."""
    if not t.is_cuda:
        raise capi.McanError("%s must be a CUDA tensor (the MCAN hot path has no CPU fallback)" % name)
    if t.device.index != _current_device():
        raise capi.McanError("%s lives on %s but the current CUDA device is %d (one process per GPU: call "
                             "torch.cuda.set_device first; nn.DataParallel is not supported)" %
                             (name, t.device, _current_device()))


def _req(t, dtype, name):
    if getattr(_tls, "depth", 0):
        return
    if not t.is_cuda:
        raise capi.McanError("%s must be a CUDA tensor (the MCAN hot path has no CPU fallback)" % name)
    if t.device.index != _current_device():
        # kernels are enqueued on the CURRENT device's current stream (one process per GPU; nn.DataParallel
        # replicas on other devices are not supported -- use torch.distributed + dp.attach)
        raise capi.McanError("%s lives on %s but the current CUDA device is %d (one process per GPU: call "
                             "torch.cuda.set_device first; nn.DataParallel is not supported)" %
                             (name, t.device, _current_device()))
    if t.dtype != dtype:
        raise capi.McanError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _req2d(t, dtype, name):
    if getattr(_tls, "depth", 0):
        return
    _req(t, dtype, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise capi.McanError("%s must be 2-D with unit inner stride" % name)


_seed_dev = None


def set_seed_tensor(t):
    """Registers a 1-element int32/uint32 CUDA tensor whose value is XOR-ed into every dropout seed
    on the device (None to clear).  Lets a captured CUDA graph draw new masks on each replay."""
    global _seed_dev
    if t is not None and not (t.is_cuda and t.numel() == 1 and t.element_size() == 4):
        raise capi.McanError("seed tensor must be a 1-element 32-bit CUDA tensor")
    _seed_dev = t


def _seed_ptr():
    return None if _seed_dev is None else _seed_dev.data_ptr()


def num_sms():
    return capi.load().mcan_num_sms()


_raw_sms = [0]


def num_sms_physical():
    """SM count of the device, regardless of set_sm_limit."""
    if _raw_sms[0] == 0:
        saved = _state["sm_limit"]
        if saved:
            capi.load().mcan_set_sm_limit(0)
        _raw_sms[0] = capi.load().mcan_num_sms()
        if saved:
            capi.load().mcan_set_sm_limit(saved)
    return _raw_sms[0]


_state = {"sm_limit": 0, "dynamic": False}


def set_sm_limit(sms):
    """Persistent kernels use at most `sms` SMs (0 = all); see include/mcan_b200.h."""
    capi.check(capi.load().mcan_set_sm_limit(int(sms)), "mcan_set_sm_limit", launch=False)
    _state["sm_limit"] = int(sms)


def set_gemm_schedule(dynamic):
    """Static (False, default) or dynamic (True) tile schedule of the GEMM; see include/mcan_b200.h."""
    capi.check(capi.load().mcan_set_gemm_schedule(1 if dynamic else 0), "mcan_set_gemm_schedule", launch=False)
    _state["dynamic"] = bool(dynamic)


class launch_config(object):
    """with launch_config(sm_limit=..., dynamic=...): GEMM launches inside use these settings."""

    def __init__(self, sm_limit=None, dynamic=None):
        self.want = (sm_limit, dynamic)

    def __enter__(self):
        self.saved = (_state["sm_limit"], _state["dynamic"])
        if self.want[0] is not None:
            set_sm_limit(self.want[0])
        if self.want[1] is not None:
            set_gemm_schedule(self.want[1])

    def __exit__(self, *exc):
        set_sm_limit(self.saved[0])
        set_gemm_schedule(self.saved[1])


def set_pdl(enabled):
    """Programmatic dependent launch between the library's kernels (default on)."""
    capi.check(capi.load().mcan_set_pdl(1 if enabled else 0), "mcan_set_pdl", launch=False)


def set_attn_impl(tcgen05):
    """Image-side attention on the tcgen05 / TMEM kernels (default) or on the mma.sync kernels."""
    capi.check(capi.load().mcan_set_attn_impl(1 if tcgen05 else 0), "mcan_set_attn_impl", launch=False)


def gemm_plan(m, n, k, *, accumulate=False, split_k=0, block_n=0, cta_group=0, sms=0):
    """The launch plan mcan_gemm uses for this shape (host logic only, no GPU needed when `sms` is given)."""
    out = (ctypes.c_int32 * 7)()
    capi.check(capi.load().mcan_gemm_plan(int(m), int(n), int(k), 1 if accumulate else 0, int(split_k), int(block_n),
                                          int(cta_group), int(sms), ctypes.cast(out, ctypes.c_void_p)), "mcan_gemm_plan",
               launch=False)
    keys = ("block_n", "cluster", "m_tiles", "n_tiles", "splits", "full_tiles", "units")
    return dict(zip(keys, (int(v) for v in out)))


_packers = {}


def _gemm_packer():
    pk = _packers.get("gemm")
    if pk is None:
        if capi.MAX_SEG != 3:
            raise capi.McanError("ops.gemm marshals exactly 3 operand segments; capi.MAX_SEG is %d" % capi.MAX_SEG)
        pk = _packers["gemm"] = capi.StructPacker(capi.GemmArgs)
    return pk


def gemm(a, b, *, a_layout=0, b_layout=0, bias=None, relu=False, dropout_p=0.0, seed=0, gate=None,
         gate_scale=1.0, resid=None, out_f32=None, out_bf16=None, out_lo=None, accumulate=False,
         split_k=0, block_n=0, cta_group=0, colsum=None):
    """D[M,N] = epilogue(sum_s A_s B_s^T) on the tcgen05 GEMM (see include/mcan_b200.h).

    a / b: bf16 tensors or equal-length lists of them (segments).  a_layout 0: [M,K], 1: [K,M];
    b_layout 0: [N,K] (nn.Linear weight), 1: [K,N].
    """
    lib = capi.load()
    a_list = list(a) if isinstance(a, (list, tuple)) else [a]
    b_list = list(b) if isinstance(b, (list, tuple)) else [b]
    if len(a_list) != len(b_list) or not 1 <= len(a_list) <= capi.MAX_SEG:
        raise capi.McanError("gemm: bad segment lists")
    for t in a_list:
        _req2d(t, _BF16, "gemm A")
    for t in b_list:
        _req2d(t, _BF16, "gemm B")
    a0, b0 = a_list[0], b_list[0]
    m, k = (a0.shape[0], a0.shape[1]) if a_layout == 0 else (a0.shape[1], a0.shape[0])
    n, kb = (b0.shape[0], b0.shape[1]) if b_layout == 0 else (b0.shape[1], b0.shape[0])
    if k != kb:
        raise capi.McanError("gemm: contraction mismatch %d vs %d" % (k, kb))
    nseg = len(a_list)
    if nseg == 1:
        pa, pb = (a0.data_ptr(), 0, 0), (b0.data_ptr(), 0, 0)
    else:
        for ta, tb in zip(a_list, b_list):
            if ta.shape != a0.shape or tb.shape != b0.shape or ta.stride(0) != a0.stride(0) or tb.stride(0) != b0.stride(0):
                raise capi.McanError("gemm: segments must share shape and leading dimension")
        pa = tuple(t.data_ptr() for t in a_list) + (0,) * (capi.MAX_SEG - nseg)
        pb = tuple(t.data_ptr() for t in b_list) + (0,) * (capi.MAX_SEG - nseg)
    p_bias = p_gate = p_resid = p_f32 = p_bf16 = p_lo = p_colsum = 0
    ldg = ldr = ldo_f32 = ldo_bf16 = 0
    if bias is not None:
        _req(bias, _F32, "gemm bias")
        p_bias = bias.data_ptr()
    if gate is not None:
        _req2d(gate, _BF16, "gemm gate")
        p_gate, ldg = gate.data_ptr(), gate.stride(0)
    if resid is not None:
        _req2d(resid, _F32, "gemm resid")
        p_resid, ldr = resid.data_ptr(), resid.stride(0)
    if out_f32 is not None:
        _req2d(out_f32, _F32, "gemm out_f32")
        p_f32, ldo_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        _req2d(out_bf16, _BF16, "gemm out_bf16")
        p_bf16, ldo_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if out_lo is not None:
        _req2d(out_lo, _BF16, "gemm out_lo")
        if out_bf16 is None or out_lo.stride(0) != out_bf16.stride(0):
            raise capi.McanError("gemm: out_lo needs out_bf16 with the same leading dimension")
        p_lo = out_lo.data_ptr()
    if colsum is not None:
        _req(colsum, _F32, "gemm colsum")
        if colsum.numel() != n or not colsum.is_contiguous():
            raise capi.McanError("gemm: colsum must be a contiguous fp32 vector of length N")
        p_colsum = colsum.data_ptr()
    # one pack_into in the field order of capi.GemmArgs (include/mcan_b200.h: mcan_gemm_args)
    args = _gemm_packer().pack(
        pa[0], pa[1], pa[2], pb[0], pb[1], pb[2], nseg, a_layout, b_layout, m, n, k, a0.stride(0), b0.stride(0),
        p_bias, 1 if relu else 0, float(dropout_p), int(seed) & 0xFFFFFFFF, _seed_ptr() or 0,
        p_gate, ldg, float(gate_scale), p_resid, ldr, p_f32, ldo_f32, p_bf16, p_lo, ldo_bf16, p_colsum,
        1 if accumulate else 0, int(split_k), int(block_n), int(cta_group), _stream() or 0)
    capi.check(lib.mcan_gemm(args), "mcan_gemm")


def gemm_grouped(problems, split_k=0, accumulate=True):
    """problems = [(a bf16 [K, M_g], b bf16 [K, N_g], out fp32 [M_g, N_g])]: out_g += a_g^T b_g for every group in ONE
    launch (the weight gradients dW = dY^T X of one layer); at most capi.MAX_GROUPS groups sharing K.
    accumulate=False: out_g = a_g^T b_g with plain stores (K not split), out_g may be uninitialised memory."""
    lib = capi.load()
    if not 1 <= len(problems) <= capi.MAX_GROUPS:
        raise capi.McanError("gemm_grouped: 1..%d problems" % capi.MAX_GROUPS)
    args = capi.GemmGroupedArgs()
    k = problems[0][0].shape[0]
    for i, (a, b, out) in enumerate(problems):
        _req2d(a, _BF16, "gemm_grouped a")
        _req2d(b, _BF16, "gemm_grouped b")
        _req2d(out, problems[0][2].dtype if problems[0][2].dtype in (_F32, _BF16) else _F32, "gemm_grouped out")
        if a.shape[0] != k or b.shape[0] != k or out.shape != (a.shape[1], b.shape[1]):
            raise capi.McanError("gemm_grouped: problem %d: shapes %s %s %s" % (i, tuple(a.shape), tuple(b.shape), tuple(out.shape)))
        g = args.g[i]
        g.a, g.b, g.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
        g.m, g.n = a.shape[1], b.shape[1]
        g.lda, g.ldb, g.ldo = a.stride(0), b.stride(0), out.stride(0)
    args.num_groups = len(problems)
    args.split_k = int(split_k)
    args.accumulate = 1 if accumulate else 0
    args.out_bf16 = 1 if problems[0][2].dtype == _BF16 else 0
    args.k = k
    args.stream = _stream()
    capi.check(lib.mcan_gemm_grouped(ctypes.byref(args)), "mcan_gemm_grouped")


def gemm_ln(a, w, *, bias, resid, ln_a2, ln_b2, eps, dropout_p=0.0, seed=0, s_f32=None, y_f32=None, y_bf16=None,
            mean=None, sigma=None):
    """y = LayerNorm(resid + dropout(a w^T + bias)) in ONE kernel (see include/mcan_b200.h, mcan_gemm_ln).
    a: bf16 [M,K]; w: bf16 [N,K] with N in {512, 1024}; outputs contiguous [M,N]."""
    lib = capi.load()
    _req2d(a, _BF16, "gemm_ln a")
    _req2d(w, _BF16, "gemm_ln w")
    _req2d(resid, _F32, "gemm_ln resid")
    m, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k or resid.shape != (m, n):
        raise capi.McanError("gemm_ln: shape mismatch")
    args = capi.GemmLnArgs()
    args.a, args.b = a.data_ptr(), w.data_ptr()
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldb = a.stride(0), w.stride(0)
    _req(bias, _F32, "gemm_ln bias")
    args.bias = bias.data_ptr()
    args.dropout_p = float(dropout_p)
    args.dropout_seed = int(seed) & 0xFFFFFFFF
    args.dropout_seed_dev = _seed_ptr()
    args.resid, args.ldr = resid.data_ptr(), resid.stride(0)
    args.ln_a2, args.ln_b2, args.eps = ln_a2.data_ptr(), ln_b2.data_ptr(), float(eps)
    for t, dt, nm in ((s_f32, _F32, "s_f32"), (y_f32, _F32, "y_f32"), (y_bf16, _BF16, "y_bf16")):
        if t is not None:
            _req(t, dt, "gemm_ln " + nm)
            if t.shape != (m, n) or not t.is_contiguous():
                raise capi.McanError("gemm_ln: %s must be contiguous [M,N]" % nm)
    args.s_f32, args.y_f32, args.y_bf16 = _ptr(s_f32), _ptr(y_f32), _ptr(y_bf16)
    args.mean, args.sigma = _ptr(mean), _ptr(sigma)
    args.stream = _stream()
    capi.check(lib.mcan_gemm_ln(ctypes.byref(args)), "mcan_gemm_ln")


def _attn_args(q, k, v, key_mask, batch, heads, sq, sk, head_dim, scale, dropout_p, seed):
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _req2d(t, _BF16, "attn " + nm)
    args = capi.AttnArgs()
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    args.ldq, args.ldk, args.ldv = q.stride(0), k.stride(0), v.stride(0)
    if key_mask is not None:
        _req(key_mask, torch.uint8, "attn key_mask")
        if key_mask.numel() != batch * sk or not key_mask.is_contiguous():
            raise capi.McanError("attn key_mask must be contiguous uint8 [batch, sk]")
        args.key_mask = key_mask.data_ptr()
    args.batch, args.heads, args.sq, args.sk, args.head_dim = batch, heads, sq, sk, head_dim
    args.scale = float(scale)
    args.dropout_p = float(dropout_p)
    args.dropout_seed = int(seed) & 0xFFFFFFFF
    args.dropout_seed_dev = _seed_ptr()
    args.stream = _stream()
    return args


def attn_fwd(q, k, v, key_mask, out, *, batch, heads, sq, sk, head_dim, scale, dropout_p=0.0, seed=0,
             q_lo=None, k_lo=None, v_lo=None, out_lo=None):
    """out[b*sq+s, h*d:(h+1)*d] = softmax(mask(Q K^T scale)) V per (batch, head).  q/k/v/out are
    bf16 [rows, >=heads*d] views (column offset already applied by slicing)."""
    lib = capi.load()
    args = _attn_args(q, k, v, key_mask, batch, heads, sq, sk, head_dim, scale, dropout_p, seed)
    _req2d(out, _BF16, "attn out")
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    if q_lo is not None:    # split precision: lo halves share the hi tensors' strides
        for hi, lo, nm in ((q, q_lo, "q_lo"), (k, k_lo, "k_lo"), (v, v_lo, "v_lo"), (out, out_lo, "out_lo")):
            _req2d(lo, _BF16, "attn " + nm)
            if lo.stride(0) != hi.stride(0):
                raise capi.McanError("attn %s must share the leading dimension of its hi tensor" % nm)
        args.q_lo, args.k_lo, args.v_lo, args.out_lo = q_lo.data_ptr(), k_lo.data_ptr(), v_lo.data_ptr(), out_lo.data_ptr()
    capi.check(lib.mcan_attn_fwd(ctypes.byref(args)), "mcan_attn_fwd")


def attn_bwd(q, k, v, key_mask, dout, dq, dk, dv, *, batch, heads, sq, sk, head_dim, scale,
             dropout_p=0.0, seed=0, dbq=None, dbk=None, dbv=None):
    lib = capi.load()
    args = capi.AttnBwdArgs()
    args.fwd = _attn_args(q, k, v, key_mask, batch, heads, sq, sk, head_dim, scale, dropout_p, seed)
    for t, nm in ((dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req2d(t, _BF16, "attn " + nm)
    args.dout, args.lddo = dout.data_ptr(), dout.stride(0)
    args.dq, args.dk, args.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    args.lddq, args.lddk, args.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    for t, nm in ((dbq, "dbq"), (dbk, "dbk"), (dbv, "dbv")):
        if t is not None:
            _req(t, _F32, "attn " + nm)
            if t.numel() != heads * head_dim or not t.is_contiguous():
                raise capi.McanError("attn %s must be a contiguous fp32 vector of heads * head_dim" % nm)
    args.dbq, args.dbk, args.dbv = _ptr(dbq), _ptr(dbk), _ptr(dbv)
    capi.check(lib.mcan_attn_bwd(ctypes.byref(args)), "mcan_attn_bwd")


def layernorm_fwd(x, a2, b2, eps, *, y_f32=None, y_bf16=None, y_lo=None, mean=None, sigma=None):
    """MCAN LayerNorm over the last dim of fp32 x [rows, h] (contiguous)."""
    lib = capi.load()
    _req(x, _F32, "layernorm x")
    if not x.is_contiguous():
        raise capi.McanError("layernorm x must be contiguous")
    h = x.shape[-1]
    rows = x.numel() // h
    capi.check(lib.mcan_layernorm_fwd(x.data_ptr(), rows, h, a2.data_ptr(), b2.data_ptr(), float(eps),
                                      _ptr(y_f32), _ptr(y_bf16), _ptr(y_lo), _ptr(mean), _ptr(sigma),
                                      _stream()), "mcan_layernorm_fwd")


def layernorm_add_fwd(x, x2, a2, b2, eps, *, s_out=None, y_f32=None, y_bf16=None, y_lo=None, mean=None, sigma=None):
    """MCAN LayerNorm of the sum x + x2 (both fp32 [rows, h], contiguous); s_out receives the sum."""
    lib = capi.load()
    _req(x, _F32, "layernorm x")
    _req(x2, _F32, "layernorm x2")
    if not (x.is_contiguous() and x2.is_contiguous()) or x.shape != x2.shape:
        raise capi.McanError("layernorm_add_fwd: x and x2 must be contiguous and of equal shape")
    h = x.shape[-1]
    rows = x.numel() // h
    capi.check(lib.mcan_layernorm_add_fwd(x.data_ptr(), x2.data_ptr(), _ptr(s_out), rows, h, a2.data_ptr(), b2.data_ptr(),
                                          float(eps), _ptr(y_f32), _ptr(y_bf16), _ptr(y_lo), _ptr(mean), _ptr(sigma),
                                          _stream()), "mcan_layernorm_add_fwd")


def rowmask_cast(x, hi, lo=None, mask=None):
    """hi (bf16 [rows, cols], row-strided) = bf16(x), lo = bf16(x - hi), mask[row] = 1 iff the fp32 row x[row] is all zero
    (net.py:135-137) -- one pass over x (fp32 [rows, cols], contiguous)."""
    lib = capi.load()
    _req(x, _F32, "rowmask_cast x")
    _req2d(hi, _BF16, "rowmask_cast hi")
    if x.dim() != 2 or not x.is_contiguous() or hi.shape != x.shape:
        raise capi.McanError("rowmask_cast: x must be contiguous 2-D and hi of the same shape")
    if lo is not None:
        _req2d(lo, _BF16, "rowmask_cast lo")
        if lo.stride(0) != hi.stride(0):
            raise capi.McanError("rowmask_cast: lo must share hi's leading dimension")
    if mask is not None:
        _req(mask, torch.uint8, "rowmask_cast mask")
        if mask.numel() != x.shape[0] or not mask.is_contiguous():
            raise capi.McanError("rowmask_cast: mask must be contiguous uint8 [rows]")
    capi.check(lib.mcan_rowmask_cast(x.data_ptr(), x.shape[0], x.shape[1], hi.data_ptr(), _ptr(lo), hi.stride(0),
                                     _ptr(mask), _stream()), "mcan_rowmask_cast")


_lstm_bar = {}


def _lstm_barrier(device):
    t = _lstm_bar.get(device)
    if t is None:
        t = _lstm_bar[device] = torch.zeros(4, dtype=torch.int32, device=device)
    return t


def embed_gather(tokens, table, x, mask=None, x_lo=None):
    """x[b * (T + 1) + t, :E] = bf16(table[tokens[b, t]]) (slot T and the pad columns zero); mask[b * T + t] = tokens == 0."""
    lib = capi.load()
    _req(tokens, torch.int64, "embed tokens")
    _req(table, _F32, "embed table")
    _req2d(x, _BF16, "embed x")
    B, T = tokens.shape
    if not (tokens.is_contiguous() and table.is_contiguous()) or x.shape[0] != B * (T + 1) or x.shape[1] != x.stride(0):
        raise capi.McanError("embed_gather: tokens / table contiguous, x = full [B * (T + 1), ld] buffer")
    if mask is not None:
        _req(mask, torch.uint8, "embed mask")
    if x_lo is not None:
        _req2d(x_lo, _BF16, "embed x_lo")
        if x_lo.shape != x.shape or x_lo.stride(0) != x.stride(0):
            raise capi.McanError("embed_gather: x_lo must have the layout of x")
    capi.check(lib.mcan_embed_gather(tokens.data_ptr(), table.data_ptr(), table.shape[0], table.shape[1], B, T, x.data_ptr(),
                                     _ptr(x_lo), x.stride(0), _ptr(mask), _stream()), "mcan_embed_gather")


def embed_scatter_add(tokens, dx, dtable):
    """dtable[tokens[b, t]] += dx[b * (T + 1) + t, :E]  (fp32; dtable zero-initialised)."""
    lib = capi.load()
    _req(tokens, torch.int64, "embed tokens")
    _req2d(dx, _F32, "embed dx")
    _req(dtable, _F32, "embed dtable")
    B, T = tokens.shape
    capi.check(lib.mcan_embed_scatter_add(tokens.data_ptr(), dx.data_ptr(), dx.stride(0), dtable.shape[0], dtable.shape[1],
                                          B, T, dtable.data_ptr(), _stream()), "mcan_embed_scatter_add")


def _lstm_args(w_hh, hbuf, batch, steps, hidden, **ptrs):
    _req2d(w_hh, _BF16, "lstm w_hh")
    if w_hh.shape != (4 * hidden, hidden) or not w_hh.is_contiguous():
        raise capi.McanError("lstm: w_hh must be contiguous bf16 [4H, H]")
    args = capi.LstmArgs()
    args.w_hh, args.hbuf = w_hh.data_ptr(), hbuf.data_ptr()
    args.batch, args.steps, args.hidden = batch, steps, hidden
    for k, v in ptrs.items():
        setattr(args, k, _ptr(v))
    args.barrier = _lstm_barrier(w_hh.device).data_ptr()
    args.stream = _stream()
    return args


def lstm_fwd(xw, w_hh, b_hh, hbuf, h_out, cbuf, gates, *, batch, steps, hidden):
    """All `steps` time steps of nn.LSTM for `batch` (<= 64) samples in one persistent launch; buffers in the
    row(b, s) = b * (steps + 1) + s layout of include/mcan_b200.h (cbuf / gates None for inference)."""
    lib = capi.load()
    _req(xw, _F32, "lstm xw")
    args = _lstm_args(w_hh, hbuf, batch, steps, hidden, xw=xw, b_hh=b_hh, h_out=h_out, cbuf=cbuf, gates=gates)
    capi.check(lib.mcan_lstm_fwd(ctypes.byref(args)), "mcan_lstm_fwd")


def lstm_bwd(dout, w_hh, hbuf, cbuf, gates, da, *, batch, steps, hidden):
    lib = capi.load()
    _req(dout, _F32, "lstm dout")
    args = _lstm_args(w_hh, hbuf, batch, steps, hidden, dout=dout, cbuf=cbuf, gates=gates, da=da)
    capi.check(lib.mcan_lstm_bwd(ctypes.byref(args)), "mcan_lstm_bwd")


_head_ws = {}


def _head_workspace(device):
    ws = _head_ws.get(device)
    if ws is None:
        ws = _head_ws[device] = torch.zeros(1032, dtype=_F32, device=device)
    return ws


def sigmoid_bce_fwd(logits, probs, target=None, loss=None):
    """probs = sigmoid(logits); with target and loss: loss[()] = BCELoss(reduction='sum')(probs, target)."""
    lib = capi.load()
    _req2d(logits, _F32, "sigmoid_bce logits")
    _req(probs, _F32, "sigmoid_bce probs")
    rows, cols = logits.shape
    if probs.shape != (rows, cols) or not probs.is_contiguous():
        raise capi.McanError("sigmoid_bce: probs must be contiguous [rows, cols]")
    ws = None
    if target is not None:
        _req(target, _F32, "sigmoid_bce target")
        if target.shape != (rows, cols) or not target.is_contiguous():
            raise capi.McanError("sigmoid_bce: target must be contiguous [rows, cols]")
    if loss is not None:
        _req(loss, _F32, "sigmoid_bce loss")
        ws = _head_workspace(logits.device)
    capi.check(lib.mcan_sigmoid_bce_fwd(logits.data_ptr(), logits.stride(0), _ptr(target), rows, cols, probs.data_ptr(),
                                        _ptr(loss), _ptr(ws), _stream()), "mcan_sigmoid_bce_fwd")


def sigmoid_bce_bwd(probs, dz, *, target=None, gout=None, gscale=None, dbias=None):
    """dz (bf16 [rows, cols] view of a padded buffer) = gradient w.r.t. the logits; see include/mcan_b200.h."""
    lib = capi.load()
    _req(probs, _F32, "sigmoid_bce probs")
    _req2d(dz, _BF16, "sigmoid_bce dz")
    rows, cols = probs.shape
    for t, nm in ((target, "target"), (gout, "gout")):
        if t is not None:
            _req(t, _F32, "sigmoid_bce " + nm)
            if t.shape != (rows, cols) or not t.is_contiguous():
                raise capi.McanError("sigmoid_bce: %s must be contiguous [rows, cols]" % nm)
    if gscale is not None:
        _req(gscale, _F32, "sigmoid_bce gscale")
    if dbias is not None:
        _req(dbias, _F32, "sigmoid_bce dbias")
    capi.check(lib.mcan_sigmoid_bce_bwd(probs.data_ptr(), _ptr(target), _ptr(gout), _ptr(gscale), rows, cols,
                                        dz.data_ptr(), dz.stride(0), _ptr(dbias), _stream()), "mcan_sigmoid_bce_bwd")


def layernorm_bwd(dy, x, mean, sigma, a2, eps, *, dx_f32=None, dx_bf16=None, dropout_p=0.0, seed=0,
                  da2=None, db2=None, dbias=None):
    lib = capi.load()
    _req(dy, _F32, "layernorm dy")
    _req(x, _F32, "layernorm x")
    if not (dy.is_contiguous() and x.is_contiguous()):
        raise capi.McanError("layernorm_bwd tensors must be contiguous")
    h = x.shape[-1]
    rows = x.numel() // h
    capi.check(lib.mcan_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), sigma.data_ptr(),
                                      a2.data_ptr(), float(eps), rows, h, _ptr(dx_f32), _ptr(dx_bf16),
                                      float(dropout_p), int(seed) & 0xFFFFFFFF, _seed_ptr(), _ptr(da2), _ptr(db2),
                                      _ptr(dbias), _stream()), "mcan_layernorm_bwd")


def attflat_pool_fwd(hmid, w2, b2, mask, x, *, batch, s, h, mlp, glimpses, att_w, pooled_f32=None,
                     pooled_bf16=None, hmid_lo=None):
    lib = capi.load()
    _req(hmid, _BF16, "attflat hmid")
    _req(x, _F32, "attflat x")
    _req(w2, _F32, "attflat w2")
    if mask is not None:
        _req(mask, torch.uint8, "attflat mask")
    capi.check(lib.mcan_attflat_pool_fwd(hmid.data_ptr(), _ptr(hmid_lo), w2.data_ptr(), b2.data_ptr(), _ptr(mask),
                                         x.data_ptr(), batch, s, h, mlp, glimpses, att_w.data_ptr(),
                                         _ptr(pooled_f32), _ptr(pooled_bf16), _stream()),
               "mcan_attflat_pool_fwd")


def attflat_pool_bwd(dpooled, pooled, hmid, w2, mask, x, att_w, *, batch, s, h, mlp, glimpses, gate_scale, dx,
                     dhmid, dw2=None, db2=None):
    lib = capi.load()
    _req(dpooled, _F32, "attflat dpooled")
    _req(pooled, _F32, "attflat pooled")
    capi.check(lib.mcan_attflat_pool_bwd(dpooled.data_ptr(), pooled.data_ptr(), hmid.data_ptr(), w2.data_ptr(), _ptr(mask),
                                         x.data_ptr(), att_w.data_ptr(), batch, s, h, mlp, glimpses,
                                         float(gate_scale), dx.data_ptr(), dhmid.data_ptr(), _ptr(dw2),
                                         _ptr(db2), _stream()), "mcan_attflat_pool_bwd")


def cast_bf16(x, hi, lo=None):
    lib = capi.load()
    _req(x, _F32, "cast x")
    _req(hi, _BF16, "cast hi")
    if not (x.is_contiguous() and hi.is_contiguous() and (lo is None or lo.is_contiguous())):
        raise capi.McanError("cast tensors must be contiguous")
    capi.check(lib.mcan_cast_bf16(x.data_ptr(), x.numel(), hi.data_ptr(), _ptr(lo), _stream()),
               "mcan_cast_bf16")


CAST_CHUNK = 4096
_CAST_F32_FLAG = 1 << 62


def build_cast_table(pairs, device):
    """pairs = [(dst, src)] or [(dst, src, dst_lo)] with fp32 contiguous src and bf16 (cast) or fp32
    (copy) contiguous dst; dst_lo (optional bf16) receives the low-order half.
    Returns (int64 table tensor on `device`, num_segments, total_chunks) for cast_multi."""
    rows, chunk = [], 0
    for item in pairs:
        dst, src = item[0], item[1]
        lo = item[2] if len(item) > 2 else None
        n = src.numel()
        if not (src.is_contiguous() and dst.is_contiguous() and dst.numel() == n and src.dtype == _F32):
            raise capi.McanError("cast table: tensors must be contiguous fp32 -> bf16/fp32 of equal size")
        flag = _CAST_F32_FLAG if dst.dtype == _F32 else 0
        rows.append([src.data_ptr(), dst.data_ptr(), n, chunk | flag, 0 if lo is None else lo.data_ptr()])
        chunk += (n + CAST_CHUNK - 1) // CAST_CHUNK
    return torch.tensor(rows, dtype=torch.int64).to(device), len(rows), chunk


def cast_multi(table, num_segments, total_chunks):
    lib = capi.load()
    capi.check(lib.mcan_cast_multi(table.data_ptr(), num_segments, total_chunks, _stream()), "mcan_cast_multi")


def gate_bf16(dy, act, scale, out):
    """out = bf16(act > 0 ? dy * scale : 0) -- backward through ReLU (+dropout) of a saved activation."""
    lib = capi.load()
    _req(dy, _F32, "gate dy")
    _req(act, _BF16, "gate act")
    _req(out, _BF16, "gate out")
    if not (dy.is_contiguous() and act.is_contiguous() and out.is_contiguous()):
        raise capi.McanError("gate tensors must be contiguous")
    capi.check(lib.mcan_gate_bf16(dy.data_ptr(), act.data_ptr(), float(scale), out.data_ptr(), dy.numel(),
                                  _stream()), "mcan_gate_bf16")


def colsum(x, out):
    """out[c] += sum_r x[r, c]  (x bf16 or fp32 2-D, out fp32)."""
    lib = capi.load()
    _req(out, _F32, "colsum out")
    if x.dtype == _BF16:
        _req2d(x, _BF16, "colsum x")
        capi.check(lib.mcan_colsum_bf16(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                                        _stream()), "mcan_colsum_bf16")
    else:
        _req2d(x, _F32, "colsum x")
        capi.check(lib.mcan_colsum_f32(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                                       _stream()), "mcan_colsum_f32")


# ---- host mirror of the device dropout hash (csrc/common.cuh), for tests --------------------
def dropout_keep_mask(numel, p, seed, device="cpu"):
    """Boolean keep-mask of the first `numel` linear element indices, identical to the kernels'."""
    import numpy as np
    idx = np.arange(numel, dtype=np.uint64)
    pair = (idx >> np.uint64(1)).astype(np.uint32)
    with np.errstate(over="ignore"):
        x = (pair * np.uint32(0x9E3779B9) + np.uint32(seed & 0xFFFFFFFF)).astype(np.uint32)
        x ^= x >> np.uint32(16)
        x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
        x ^= x >> np.uint32(15)
        x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
        x ^= x >> np.uint32(16)
    u16 = np.where((idx & np.uint64(1)) == 1, x >> np.uint32(16), x & np.uint32(0xFFFF))
    thr = min(max(int(p * 65536.0 + 0.5), 0), 65535)
    return torch.from_numpy((u16 >= thr)).to(device)
