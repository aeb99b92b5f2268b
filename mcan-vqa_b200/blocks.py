"""Forward / backward of the MCAN blocks as chains of libmcan_b200 kernels.

This is the host side of the hot path: it owns no arithmetic, only the order of kernel
launches, the activation buffers (torch tensors used as plain device memory) and the
bookkeeping autograd needs.  Reference semantics (all in /root/reference):

    MHAtt  core/model/mca.py:30-78      FFN   core/model/mca.py:85-98 (+ net_utils.py:11-45)
    SA     core/model/mca.py:118-127    SGA   core/model/mca.py:150-164
    MCA_ED core/model/mca.py:178-186    AttFlat core/model/net.py:38-55
    LayerNorm core/model/net_utils.py:56-60

Data layout in HBM.  Every activation is a row-major [rows, features] matrix, rows =
batch*sequence.  The residual stream is kept twice: fp32 (residual adds, LayerNorm input) and
bf16 (the A operand of the next tcgen05 GEMM), both written by the LayerNorm kernel.  Q/K/V
live in one bf16 [rows, 3H] buffer written by one fused GEMM; heads are column slices, so the
reference's view/transpose/contiguous copies (mca.py:33-59) do not exist.  For MCA_ED the K/V
projections of the final encoder output for ALL decoder layers are one GEMM into a
[rows_q, L*2H] buffer.  Weights: fp32 masters (the nn.Parameters the optimiser updates) and
cached bf16 copies stacked per GEMM ([q|k|v] = [3H,H]); dgrad and wgrad read the same bf16
buffers MN-major, no transposed copies.  Dropout masks are never materialised.

Per sub-layer launch chain (forward):
    self-attention: GEMM[qkv] -> attention -> GEMM[merge + bias + dropout + residual] -> LayerNorm
    FFN:            GEMM[fc + bias + ReLU + dropout] -> GEMM[out + bias + dropout + residual] -> LayerNorm
"""
import math
import os

import torch

from . import ops

_BF16 = torch.bfloat16
_F32 = torch.float32

# The bf16 operand copies of the weights are re-cast on EVERY training forward (grad enabled):
# optimizer updates cannot be observed reliably from here (fused multi-tensor optimizers do not
# bump tensor version counters, and under CUDA-graph replay no Python runs at all).  Forwards
# without grad re-cast only when a master changed (version / storage) or a training forward
# happened since the last refresh.  Cost: one multi-tensor cast per step (refresh_params).
ALWAYS_RECAST = True

# Host-logic dry run (tests/test_hostlogic_cpu.py only): the C ABI is replaced by a stub that launches nothing, so the
# whole launch chain -- buffer shapes, arenas, grouping, autograd plumbing -- runs on CPU tensors without a GPU.  The
# product never sets this; with it unset every entry point still refuses CPU tensors (no fallback).
DRY_RUN = False


# "bf16": bf16 GEMM/attention operands, fp32 accumulation (training and default inference).
# "fp32": split-precision INFERENCE -- every operand is carried as bf16 hi + lo and every product is
#         evaluated as hi*hi + hi*lo + lo*hi on the tensor cores (~2e-5 relative, i.e. fp32-grade
#         logits and identical top-1 answers); applies to forwards without grad, training stays bf16.
PRECISION = os.environ.get("MCAN_PRECISION", "bf16")


def set_precision(mode):
    global PRECISION
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    PRECISION = mode


def _mix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


_seed_counter = [0]


class Runtime(object):
    """Per-forward settings: dropout probability (0 in eval) and the dropout seed stream."""

    def __init__(self, training, dropout_rate):
        self.p = float(dropout_rate) if training else 0.0
        _seed_counter[0] += 1
        self.base = _mix32((torch.initial_seed() & 0xFFFFFFFF) ^ (_seed_counter[0] * 0x9E3779B9))
        self.n = 0
        self.bufs = []      # loose fp32 gradient buffers produced since the last drain (for dp.py)
        # (a Runtime is created by the module-level wrapper, OUTSIDE autograd.Function.forward, where grad mode is
        #  always off: this is the caller's grad mode)
        self.grad = torch.is_grad_enabled()
        self.split = (PRECISION == "fp32") and not self.grad
        self.arena = None   # one zero-initialised flat buffer all gradients of a backward are carved from
        self.arena_w = None  # uninitialised flat buffer for the weight gradients that grouped launches write
        self.scratch = None  # zero-initialised pool for the split-K outputs of the backward chain
        self.scratch_off = 0
        self.arena_w_off = 0
        self.arena_w_mark = 0
        self.arena_off = 0
        self.arena_mark = 0
        self.deferred = None   # list of (fn, args, kwargs) while weight-gradient work is being deferred
        self.group = None      # list of (dy, x, out) weight-gradient GEMMs of the layer being back-propagated

    def wgrad(self, fn, *args, **kw):
        """Weight-gradient work (wgrad GEMMs, bias column sums): nothing downstream in the backward
        chain depends on it, so MCA_ED's backward may defer it to a second stream (mca_ed_bwd)."""
        if self.deferred is not None:
            self.deferred.append((fn, args, kw))
        else:
            fn(*args, **kw)

    def wgrad_gemm(self, dy, x, out):
        """out[M,N] (fp32, zero-initialised) += dy[K,M]^T x[K,N] -- a weight gradient.  While a layer's backward is
        collecting (begin_group), the GEMM is queued and launched together with the layer's other weight gradients
        as one grouped kernel (end_group)."""
        if self.deferred is not None:
            self.deferred.append((ops.gemm, (dy, x), dict(a_layout=1, b_layout=1, out_f32=out, accumulate=True)))
        elif self.group is not None:
            self.group.append((dy, x, out, self.in_store_arena(out)))
        else:
            ops.gemm(dy, x, a_layout=1, b_layout=1, out_f32=out, accumulate=True)

    def begin_group(self):
        if GROUP_WGRADS and self.deferred is None:
            self.group = []

    def end_group(self):
        """Launches the queued weight-gradient GEMMs: one grouped launch per contraction length (<= 8 problems each).
        Gradients that live in the store arena are WRITTEN (K not split), the others accumulated into zeroed memory."""
        todo, self.group = self.group, None
        if not todo:
            return
        by_key = {}
        for dy, x, out, store in todo:
            by_key.setdefault((dy.shape[0], store), []).append((dy, x, out))
        for (_, store), items in by_key.items():
            for i in range(0, len(items), ops.capi.MAX_GROUPS):
                chunk = items[i:i + ops.capi.MAX_GROUPS]
                if len(chunk) == 1 and chunk[0][2].dtype == _BF16:
                    ops.gemm(chunk[0][0], chunk[0][1], a_layout=1, b_layout=1, out_bf16=chunk[0][2])
                elif len(chunk) == 1:
                    ops.gemm(chunk[0][0], chunk[0][1], a_layout=1, b_layout=1, out_f32=chunk[0][2], accumulate=not store)
                else:
                    ops.gemm_grouped(chunk, accumulate=not store)

    def in_store_arena(self, t):
        a = self.arena_w
        return a is not None and a.data_ptr() <= t.data_ptr() < a.data_ptr() + a.numel() * a.element_size()

    def empty_w(self, n, device):
        """Uninitialised storage for a weight gradient that a grouped launch will WRITE; None when not available
        (no layer group open, arena exhausted): the caller then takes zeroed memory and accumulates."""
        n4 = (n + 7) // 8 * 8          # 16-byte aligned carve-outs for fp32 and bf16
        a = self.arena_w
        if self.group is None or a is None or a.device != device or self.arena_w_off + n4 > a.numel():
            return None
        v = a[self.arena_w_off:self.arena_w_off + n]
        self.arena_w_off += n4
        return v

    def scratch_zeros(self, m, n, device):
        """Zero-initialised fp32 [m, n] for a split-K output (an activation gradient, not a parameter gradient)."""
        need = (m * n + 3) // 4 * 4
        p = self.scratch
        if p is None or p.device != device or self.scratch_off + need > p.numel():
            return torch.zeros((m, n), dtype=_F32, device=device)
        v = p[self.scratch_off:self.scratch_off + m * n].view(m, n)
        self.scratch_off += need
        return v

    def use_arena(self, numel, device, store_numel=0, scratch_numel=0):
        """One memset instead of one per gradient tensor; contiguous per-layer slices for dp.py.  store_numel > 0:
        that many elements are an UNINITIALISED second arena for the weight gradients grouped launches write."""
        self.arena = torch.zeros(numel, dtype=_F32, device=device)
        self.arena_off = 0
        self.arena_mark = 0
        self.arena_w = None
        if store_numel > 0 and STORE_WGRADS:
            self.arena_w = torch.empty(store_numel, dtype=_BF16 if WGRAD_BF16 else _F32, device=device)
        self.arena_w_off = 0
        self.arena_w_mark = 0
        self.scratch = torch.zeros(scratch_numel, dtype=_F32, device=device) if scratch_numel > 0 else None
        self.scratch_off = 0

    def zeros(self, n, device):
        n4 = (n + 3) // 4 * 4          # keep every carve-out 16-byte aligned
        if self.arena is not None and self.arena.device == device and self.arena_off + n4 <= self.arena.numel():
            v = self.arena[self.arena_off:self.arena_off + n]
            self.arena_off += n4
            return v
        t = torch.zeros(n, dtype=_F32, device=device)
        self.bufs.append(t)
        return t

    def drain(self):
        """Gradient storage produced since the previous drain, as a list of flat tensors."""
        out = list(self.bufs)
        self.bufs = []
        if self.arena is not None and self.arena_off > self.arena_mark:
            out.append(self.arena[self.arena_mark:self.arena_off])
            self.arena_mark = self.arena_off
        if self.arena_w is not None and self.arena_w_off > self.arena_w_mark:
            out.append(self.arena_w[self.arena_w_mark:self.arena_w_off])
            self.arena_w_mark = self.arena_w_off
        return out

    def seed(self):
        self.n += 1
        return _mix32(self.base + self.n * 0x85EBCA6B)

    @property
    def keep_scale(self):
        return 1.0 / (1.0 - self.p) if self.p > 0 else 1.0


class Act(object):
    """An activation in the formats the kernels consume: fp32 (residual stream) and bf16 (GEMM operand)."""
    __slots__ = ("f32", "bf", "lo")

    def __init__(self, f32=None, bf=None, lo=None):
        self.f32 = f32
        self.bf = bf
        self.lo = lo        # bf16(x - bf16(x)), split-precision mode only


def act_from_f32(x2d, split=False):
    bf = torch.empty(x2d.shape, dtype=_BF16, device=x2d.device)
    lo = torch.empty(x2d.shape, dtype=_BF16, device=x2d.device) if split else None
    ops.cast_bf16(x2d, bf, lo)
    return Act(x2d, bf, lo)


def _mm(rt, a, lp, i0, i1, out_bf16=None, out_lo=None, **kw):
    """GEMM of activation `a` (Act) with stacked members [i0,i1) of `lp`; 3 operand segments in
    split-precision mode."""
    w, b = lp.rows(i0, i1)
    if rt.split:
        wl = lp.rows_lo(i0, i1)
        ops.gemm([a.bf, a.bf, a.lo], [w, wl, w], bias=b, out_bf16=out_bf16, out_lo=out_lo, **kw)
    else:
        ops.gemm(a.bf, w, bias=b, out_bf16=out_bf16, **kw)


# The weight-gradient GEMMs of one SA / SGA layer feed nothing in the backward chain: they are queued while the layer's
# chain is enqueued and run as ONE grouped launch per contraction length (mcan_gemm_grouped).  Alone, each of them
# pays the fixed ~8-10 us of a launch and fills the 74 CTA pairs only partly (1024 x 1024 x 6400: 16 tiles; the
# 896-row encoder wgrads: 12-20 us for 2-7 GFLOP each); together their tiles form full waves.  MCAN_GROUP_WGRADS=0: off.
GROUP_WGRADS = os.environ.get("MCAN_GROUP_WGRADS", "1") != "0"
# With K unsplit every element of a grouped weight gradient is produced by exactly one work unit, so it is WRITTEN
# into uninitialised memory instead of accumulated into a zero-filled arena: for MCAN-large that removes an 0.7 GB
# memset per step and the read-modify-write of the same bytes by the fp32 atomics.  MCAN_STORE_WGRADS=0: off.
STORE_WGRADS = os.environ.get("MCAN_STORE_WGRADS", "1") != "0"
# Data parallel with the bf16 gradient exchange (set by train.Trainer): the grouped wgrad launches write their weight
# gradients as bf16 straight into the buffer the all-reduce works on -- no fp32 gradient, no cast pass (0.8 GB read +
# 0.4 GB written per step) for 95 % of the parameters.  Those parameters get no .grad: the bucket-wise fused optimiser
# receives (parameter, reduced bf16 gradient) pairs from dp.GradSync.
WGRAD_BF16 = False


# Bias gradients of the Q / K / V projections out of the attention backward kernel (column sums of dQ / dK / dV reduced
# per warp in shared memory, one global atomic per column and CTA) instead of column-sum launches over the gradient
# buffers.  Implemented, parity-tested and measured SLOWER: 8.01 ms/step with vs 7.94 ms without (18 fewer launches,
# but the backward attention kernel sits at 250 registers and one CTA per SM; the extra shuffles and shared-memory
# traffic cost more than 18 column-sum launches of 8 us that hide in the gaps of the chain).  Off by default.
ATTN_BIAS_GRADS = os.environ.get("MCAN_ATTN_BIAS_GRADS", "0") != "0"


# Short-M GEMMs with a long contraction (the 896-row question side: FFN2 forward, the dgrads of
# FFN1 / QKV / the batched K,V projection) fill only 16-32 CTA pairs and run 64-192 k-blocks each.
# Their epilogues (bias, dropout, residual) are linear in the accumulator, so they run split-K
# (red.global.add into a zeroed fp32 output, K split 0 adds bias + residual) on all SMs instead.
SPLITK_MAX_ROWS = 1024
SPLITK_MIN_K = 2048
# The forward GEMMs of this class (encoder FFN2, the 64-row answer projection) stay unsplit by default: an A/B on one
# box gave 7.98 ms/step without vs 8.01 ms with (the zero-fill of the output eats the gain), and an unsplit forward is
# bit-reproducible run to run.  MCAN_SPLITK_FWD=1 switches them to split-K when gradients are being recorded.
SPLITK_FWD = os.environ.get("MCAN_SPLITK_FWD", "0") != "0"
# MCAN_SPLITK_HEAD=1: only the 64-row answer projection (64 x 3129 x 2048: 34.8 us unsplit, 33.6 us as 125 split-K units --
# no gain, so off: the forward stays free of fp32 atomics)
SPLITK_HEAD = os.environ.get("MCAN_SPLITK_HEAD", "0") != "0"


def _resid_gemm(a, w, M, N, K, dev, rt=None, **kw):
    """out_f32[M,N] = epilogue(a w^T) for an epilogue without ReLU / bf16 output; split-K when short and deep.
    Used by the backward chain only (the forward's FFN2 applies the same rule when grad is enabled).  With `rt` the
    zero-initialised output comes out of the Runtime's scratch pool (one memset per backward pass instead of one each)."""
    if SPLITK_MIN_K > 0 and M <= SPLITK_MAX_ROWS and K >= SPLITK_MIN_K:
        out = rt.scratch_zeros(M, N, dev) if rt is not None else torch.zeros((M, N), dtype=_F32, device=dev)
        ops.gemm(a, w, out_f32=out, accumulate=True, **kw)
    else:
        out = _empty(M, N, _F32, dev)
        ops.gemm(a, w, out_f32=out, **kw)
    return out


def _empty(rows, cols, dtype, device):
    return torch.empty((rows, cols), dtype=dtype, device=device)


def _bf_padded(rows, cols, device):
    """bf16 [rows, cols] view whose leading dimension is a multiple of 8 (TMA stride alignment)."""
    ld = (cols + 7) // 8 * 8
    buf = torch.zeros((rows, ld), dtype=_BF16, device=device) if ld != cols else \
        torch.empty((rows, ld), dtype=_BF16, device=device)
    return buf[:, :cols]


class LinearParams(object):
    """bf16 GEMM-operand copy of one or more nn.Linear layers stacked along the output dim.

    `pairs` = [(weight Parameter [n_i, k], bias Parameter [n_i]), ...].  The copy is refreshed
    when a master's version counter or storage changes (optimizer step, load_state_dict, .cuda()).
    """

    def __init__(self, pairs, pad=8):
        self.pairs = list(pairs)
        self.pad = pad          # the copy's leading dimension is k rounded up to a multiple of `pad` (zero-filled columns)
        self.w_full = None      # the copy with its pad columns: [n, ld]
        self.sizes = [w.shape[0] for w, _ in self.pairs]
        self.n = sum(self.sizes)
        self.k = self.pairs[0][0].shape[1]
        self.w = None
        self.w_lo = None
        self.b = None
        self._stamp = None
        self._dirty = False
        self.managed = None     # optim.FusedAdamW that re-emits the operand copy with every parameter update
        self._lo_epoch = -1
        self._scope_epoch = -1  # id of the refresh_scope that last refreshed this copy

    def shadow_items(self):
        """[(dst, master Parameter)] an optimiser has to keep in sync, or None when the copy cannot be
        written by a flat kernel (padded leading dimension)."""
        self._ensure_storage()
        if self.w.stride(0) != self.k:
            return None
        out, r = [], 0
        for (w, b), n in zip(self.pairs, self.sizes):
            out.append((self.w[r:r + n], w))
            if len(self.pairs) > 1:
                out.append((self.b[r:r + n], b))
            r += n
        return out

    def _current_stamp(self):
        return tuple((w.data_ptr(), w._version, b.data_ptr(), b._version) for w, b in self.pairs)

    def _ensure_storage(self):
        dev = self.pairs[0][0].device
        if self.w is None or self.w.device != dev:
            ld = (self.k + self.pad - 1) // self.pad * self.pad
            self.w_full = torch.zeros((self.n, ld), dtype=_BF16, device=dev)
            self.w = self.w_full[:, :self.k]
            self.b = torch.empty((self.n,), dtype=_F32, device=dev) if len(self.pairs) > 1 else None
            self.w_lo = None
            self._stamp = None

    def pending(self, force=False, need_lo=False):
        """(dst, src[, dst_lo]) copy items needed to bring the bf16 operand copy up to date ([] if current)."""
        if (not force and _refresh_scope[0] > 0 and self._scope_epoch == _scope_epoch[0] and self.w is not None
                and (not need_lo or self.w_lo is not None)):
            return []       # the enclosing refresh_scope brought this copy up to date when it was entered
        self._ensure_storage()
        if need_lo and self.w_lo is None:
            self.w_lo = torch.zeros((self.n, self.w.stride(0)), dtype=_BF16, device=self.w.device)[:, :self.k]
            force = True
        stamp = self._current_stamp()
        if self.managed is not None and stamp == self._stamp:
            # the optimiser rewrites the bf16 copy together with the masters; only the low-order
            # halves (split-precision inference) are refreshed here, once per optimiser epoch
            if self.w_lo is None or not need_lo or (not force and self._lo_epoch == self.managed.epoch):
                return []
            self._lo_epoch = self.managed.epoch
            force = True
        # inside a refresh_scope the scope entry already brought every copy up to date
        dirty = self._dirty and _refresh_scope[0] == 0
        if not force and not dirty and stamp == self._stamp:
            return []
        self._dirty = bool(force) and self.managed is None   # a training forward: the optimizer may change the masters afterwards
        out = []
        r = 0
        for (w, b), n in zip(self.pairs, self.sizes):
            if self.w_lo is not None:
                out.append((self.w[r:r + n], w.detach(), self.w_lo[r:r + n]))
            else:
                out.append((self.w[r:r + n], w.detach()))
            if len(self.pairs) > 1:
                out.append((self.b[r:r + n], b.detach()))
            r += n
        if len(self.pairs) == 1:
            self.b = self.pairs[0][1].detach()      # single layer: use the fp32 master bias in place
        self._stamp = stamp
        return out

    def get(self, force=False, need_lo=False):
        refresh_params([self], force, need_lo)
        return self

    def rows_lo(self, i0, i1):
        r0 = sum(self.sizes[:i0])
        r1 = sum(self.sizes[:i1])
        return self.w_lo[r0:r1]

    def rows(self, i0, i1):
        """(weight rows, bias) of stacked members i0..i1-1."""
        r0 = sum(self.sizes[:i0])
        r1 = sum(self.sizes[:i1])
        return self.w[r0:r1], self.b[r0:r1]


class GradBuf(object):
    """Zero-initialised fp32 gradient storage for members [i0, i1) of a stacked LinearParams
    (wgrad accumulates with atomics, so the buffer must start at zero)."""

    def __init__(self, rt, lp, i0=0, i1=None):
        i1 = len(lp.sizes) if i1 is None else i1
        self.lp, self.i0, self.i1 = lp, i0, i1
        self.sizes = lp.sizes[i0:i1]
        n = sum(self.sizes)
        w = rt.empty_w(n * lp.k, lp.w.device)
        if w is None:
            flat = rt.zeros(n * lp.k + n, lp.w.device)
            self.w = flat[: n * lp.k].view(n, lp.k)
            self.b = flat[n * lp.k:]
        else:       # written (not accumulated) by the layer's grouped wgrad launch; the bias gradient needs zeros
            self.w = w.view(n, lp.k)
            self.b = rt.zeros(n, lp.w.device)

    def _range(self, j0, j1):
        r0 = sum(self.lp.sizes[self.i0:j0])
        return r0, r0 + sum(self.lp.sizes[j0:j1])

    def rows_w(self, j0, j1):
        r0, r1 = self._range(j0, j1)
        return self.w[r0:r1]

    def rows_b(self, j0, j1):
        r0, r1 = self._range(j0, j1)
        return self.b[r0:r1]

    def per_param(self):
        """[(dW_i, db_i)] views for members i0..i1-1."""
        out = []
        r = 0
        for n in self.sizes:
            out.append((self.w[r:r + n], self.b[r:r + n]))
            r += n
        return out


class Bag(object):
    pass


_refresh_scope = [0]     # > 0 while an enclosing module already refreshed every operand copy
_scope_epoch = [0]       # counts outermost refresh_scopes; LinearParams._scope_epoch == this: refreshed by the current one


def _force(rt_training):
    return ALWAYS_RECAST and rt_training and _refresh_scope[0] == 0


def refresh_params(lps, force=False, need_lo=False):
    """Brings the bf16 operand copies of many layers up to date with ONE multi-tensor cast kernel
    (fp32 -> bf16 [+ lo] of the weights, concatenation of stacked biases)."""
    todo = []
    for lp in lps:
        todo += lp.pending(force, need_lo)
    if not todo:
        return

    def ok(item):
        return all(t.is_contiguous() for t in item) and item[1].dtype == _F32

    fast = [it for it in todo if ok(it)]
    slow = [it for it in todo if not ok(it)]
    if fast:
        key = tuple(tuple(t.data_ptr() for t in it) + (it[1].numel(),) for it in fast)
        entry = _cast_tables.get(key)
        if entry is None:
            if len(_cast_tables) > 64:
                _cast_tables.clear()
            entry = ops.build_cast_table(fast, fast[0][0].device)
            _cast_tables[key] = entry
        ops.cast_multi(*entry)
    for it in slow:    # padded leading dimensions (k % 8 != 0, e.g. the 300-wide LSTM input weight): strided cast
        if it[0].dim() == 2 and it[1].dim() == 2 and it[1].is_contiguous() and it[0].dtype == _BF16 and it[0].stride(1) == 1:
            ops.rowmask_cast(it[1], it[0], it[2] if len(it) > 2 else None)
        else:
            it[0].copy_(it[1])
            if len(it) > 2:
                it[2].copy_(it[1] - it[0].float())


_cast_tables = {}


class refresh_scope(object):
    """with refresh_scope(lps, training): all operand copies are refreshed once, up front."""

    def __init__(self, lps, training):
        self.lps, self.training = lps, training

    def __enter__(self):
        if _refresh_scope[0] == 0:
            need_lo = (PRECISION == "fp32") and not torch.is_grad_enabled()
            _scope_epoch[0] += 1
            refresh_params(self.lps, ALWAYS_RECAST and self.training, need_lo)
            for lp in self.lps:
                lp._scope_epoch = _scope_epoch[0]
        _refresh_scope[0] += 1

    def __exit__(self, *exc):
        _refresh_scope[0] -= 1


# ------------------------------------------------------------------------------------------
# GEMM with residual add + LayerNorm fused into the epilogue (mcan_gemm_ln, csrc/gemm_ln.cu)
# ------------------------------------------------------------------------------------------
# "1": every eligible sub-layer output; "0" (default): GEMM -> s -> LayerNorm kernel.  One cluster owns whole rows,
# so the hidden size must be 512 or 1024; split-precision inference and split-K GEMMs keep the unfused chain.
# Opt-in because at batch 64 it is SLOWER (one wave of 25 clusters: the two-pass epilogue is fully exposed --
# measurements in csrc/gemm_ln.cu and DESIGN.md): 8.46 / 8.85 ms per step with it vs 8.28 ms without.
FUSE_LN = os.environ.get("MCAN_FUSE_LN", "0") != "0"
FUSE_LN_MIN_ROWS = int(os.environ.get("MCAN_FUSE_LN_MIN_ROWS", "1"))


def _can_fuse_ln(rt, rows, n, lp):
    return (FUSE_LN and not rt.split and n in (512, 1024) and rows >= FUSE_LN_MIN_ROWS and lp.k % 8 == 0
            and lp.w.stride(0) == lp.k)


def gemm_ln_fwd(a_bf, lp, norm, resid, p, seed):
    """LN(resid + dropout(a W^T + b)) -> (Act(y_f32, y_bf), s_f32, mean, sigma) in one launch."""
    dev = a_bf.device
    M, N = a_bf.shape[0], lp.n
    s = _empty(M, N, _F32, dev)
    y32 = _empty(M, N, _F32, dev)
    ybf = _empty(M, N, _BF16, dev)
    mean = torch.empty(M, dtype=_F32, device=dev)
    sigma = torch.empty(M, dtype=_F32, device=dev)
    ops.gemm_ln(a_bf, lp.w, bias=lp.b, resid=resid, ln_a2=norm.a_2.detach(), ln_b2=norm.b_2.detach(), eps=norm.eps,
                dropout_p=p, seed=seed, s_f32=s, y_f32=y32, y_bf16=ybf, mean=mean, sigma=sigma)
    return Act(y32, ybf), s, mean, sigma


# ------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------
def ln_fwd(norm, s_f32, want_bf=True, split=False):
    """s_f32 [rows, h] fp32 -> (Act(y_f32, y_bf[, y_lo]), mean, sigma)."""
    rows, h = s_f32.shape
    dev = s_f32.device
    y32 = _empty(rows, h, _F32, dev)
    ybf = _empty(rows, h, _BF16, dev) if want_bf else None
    ylo = _empty(rows, h, _BF16, dev) if (want_bf and split) else None
    mean = torch.empty(rows, dtype=_F32, device=dev)
    sigma = torch.empty(rows, dtype=_F32, device=dev)
    ops.layernorm_fwd(s_f32, norm.a_2.detach(), norm.b_2.detach(), norm.eps, y_f32=y32, y_bf16=ybf, y_lo=ylo,
                      mean=mean, sigma=sigma)
    return Act(y32, ybf, ylo), mean, sigma


def ln_bwd(rt, norm, dy, s_f32, mean, sigma, p=0.0, seed=0, want_bf=True, dbias=None):
    """Returns (dx_f32, dx_bf gated by the dropout mask of the producing GEMM, da2, db2)."""
    rows, h = s_f32.shape
    dev = s_f32.device
    dx = _empty(rows, h, _F32, dev)
    dxbf = _empty(rows, h, _BF16, dev) if want_bf else None
    dab = rt.zeros(2 * h, dev)
    ops.layernorm_bwd(dy, s_f32, mean, sigma, norm.a_2.detach(), norm.eps, dx_f32=dx, dx_bf16=dxbf,
                      dropout_p=p, seed=seed, da2=dab[:h], db2=dab[h:], dbias=dbias)
    return dx, dxbf, dab[:h], dab[h:]


# ------------------------------------------------------------------------------------------
# attention sub-layer:  LN(x + dropout(merge(att(...))))   (mca.py:119-121, 152-158)
# With norm=None it is the bare MHAtt module (merge output, fp32, no residual).
# ------------------------------------------------------------------------------------------
def att_fwd(rt, mh, x, B, Sq, kv_src=None, Sk=None, key_mask=None, kv=None, norm=None, v_src=None):
    """x: Act of the query-side input [B*Sq, H].
    kv_src None and kv None -> self-attention (fused QKV GEMM).
    kv_src Act             -> K,V projected from kv_src (fused KV GEMM); v_src Act additionally
                              gives V its own input (fully general MHAtt(v,k,q)).
    kv (k_view, v_view)    -> K,V already projected (MCA_ED cross-layer batch).
    Returns (Act or fp32 tensor, ctx)."""
    H = mh.hidden_size
    heads, d = mh.multi_head, mh.head_dim
    dev = x.bf.device
    M = B * Sq
    c = Bag()
    c.B, c.Sq, c.mask = B, Sq, key_mask
    lp = mh.lp_qkv().get(_force(rt.p > 0 or rt.grad), rt.split)
    c.lp = lp
    c.x_bf = x.bf
    c.mode = "self" if (kv_src is None and kv is None) else ("kv" if kv is not None else "cross")
    sp = rt.split
    q_lo = k_lo = v_lo = None
    if c.mode == "self":
        Sk = Sq
        qkv = _empty(M, 3 * H, _BF16, dev)
        qkv_lo = _empty(M, 3 * H, _BF16, dev) if sp else None
        _mm(rt, x, lp, 0, 3, out_bf16=qkv, out_lo=qkv_lo)
        q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
        if sp:
            q_lo, k_lo, v_lo = qkv_lo[:, :H], qkv_lo[:, H:2 * H], qkv_lo[:, 2 * H:]
    else:
        q = _empty(M, H, _BF16, dev)
        q_lo = _empty(M, H, _BF16, dev) if sp else None
        _mm(rt, x, lp, 0, 1, out_bf16=q, out_lo=q_lo)
        if c.mode == "kv":
            if sp:
                k, v, k_lo, v_lo = kv
            else:
                k, v = kv[0], kv[1]
        else:
            kvb = _empty(B * Sk, 2 * H, _BF16, dev)
            kvl = _empty(B * Sk, 2 * H, _BF16, dev) if sp else None
            k, v = kvb[:, :H], kvb[:, H:]
            if sp:
                k_lo, v_lo = kvl[:, :H], kvl[:, H:]
            if v_src is None:
                _mm(rt, kv_src, lp, 1, 3, out_bf16=kvb, out_lo=kvl)
            else:
                _mm(rt, kv_src, lp, 1, 2, out_bf16=k, out_lo=k_lo)
                _mm(rt, v_src, lp, 2, 3, out_bf16=v, out_lo=v_lo)
                c.v_bf = v_src.bf
            c.kv_bf = kv_src.bf
    c.Sk = Sk
    c.q, c.k, c.v = q, k, v
    att = _empty(M, H, _BF16, dev)
    att_lo = _empty(M, H, _BF16, dev) if sp else None
    c.seed_att = rt.seed()
    ops.attn_fwd(q, k, v, key_mask, att, batch=B, heads=heads, sq=Sq, sk=Sk, head_dim=d,
                 scale=1.0 / math.sqrt(d), dropout_p=rt.p, seed=c.seed_att,
                 q_lo=q_lo, k_lo=k_lo, v_lo=v_lo, out_lo=att_lo)
    c.att = att
    lpm = mh.lp_merge().get(_force(rt.p > 0 or rt.grad), rt.split)
    c.lpm = lpm
    atta = Act(None, att, att_lo)
    if norm is None:
        s = _empty(M, H, _F32, dev)
        _mm(rt, atta, lpm, 0, 1, out_f32=s)
        return s, c
    c.seed_out = rt.seed()
    if _can_fuse_ln(rt, M, H, lpm):
        out, c.s, c.mean, c.sigma = gemm_ln_fwd(att, lpm, norm, x.f32, rt.p, c.seed_out)
        return out, c
    s = _empty(M, H, _F32, dev)
    _mm(rt, atta, lpm, 0, 1, dropout_p=rt.p, seed=c.seed_out, resid=x.f32, out_f32=s)
    out, c.mean, c.sigma = ln_fwd(norm, s, split=rt.split)
    c.s = s
    return out, c


def att_bwd(rt, mh, c, dout, norm=None, dkv=None, need_dx=True):
    """dout: fp32 grad of the sub-layer output (or of the bare merge output when norm is None).
    Returns (dx_f32, dkv_src_f32 or None, dv_src_f32 or None, grads) where grads maps
    parameter -> gradient tensor.  With dkv=(dk_view, dv_view) the K/V gradients are written
    there (MCA_ED cross-layer batch) instead of being back-projected here."""
    H = mh.hidden_size
    heads, d = mh.multi_head, mh.head_dim
    dev = dout.device
    B, Sq, Sk = c.B, c.Sq, c.Sk
    M = B * Sq
    grads = {}
    gm = GradBuf(rt, c.lpm)
    if norm is None:
        ds_f32 = None
        ds_bf = _empty(M, H, _BF16, dev)
        ops.cast_bf16(dout.contiguous(), ds_bf)
        rt.wgrad(ops.colsum, dout, gm.b)
    else:
        ds_f32, ds_bf, da2, db2 = ln_bwd(rt, norm, dout, c.s, c.mean, c.sigma, p=rt.p, seed=c.seed_out, dbias=gm.b)
        grads[norm.a_2], grads[norm.b_2] = da2, db2
    # merge linear: wgrad + dgrad
    rt.wgrad_gemm(ds_bf, c.att, gm.w)
    datt = _empty(M, H, _BF16, dev)
    ops.gemm(ds_bf, c.lpm.w, b_layout=1, out_bf16=datt)
    (wm, bm), = gm.per_param()
    grads[mh.linear_merge.weight], grads[mh.linear_merge.bias] = wm, bm
    # attention backward
    lp = c.lp
    g = GradBuf(rt, lp, 0, 1 if c.mode == "kv" else 3)
    dkv_src = dv_src = None
    if c.mode == "self":
        dqkv = _empty(M, 3 * H, _BF16, dev)
        dq, dk, dv = dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:]
    else:
        dq = _empty(M, H, _BF16, dev)
        if dkv is not None:
            dk, dv = dkv
        else:
            dkvb = _empty(B * Sk, 2 * H, _BF16, dev)
            dk, dv = dkvb[:, :H], dkvb[:, H:]
    # the bias gradients of the projections this block owns (column sums of dq / dk / dv) come out of the attention
    # backward kernel itself; in "kv" mode K and V belong to the caller's batched projection
    own_kv = c.mode != "kv"
    if ATTN_BIAS_GRADS:
        ops.attn_bwd(c.q, c.k, c.v, c.mask, datt, dq, dk, dv, batch=B, heads=heads, sq=Sq, sk=Sk, head_dim=d,
                     scale=1.0 / math.sqrt(d), dropout_p=rt.p, seed=c.seed_att, dbq=g.rows_b(0, 1),
                     dbk=g.rows_b(1, 2) if own_kv else None, dbv=g.rows_b(2, 3) if own_kv else None)
    else:
        ops.attn_bwd(c.q, c.k, c.v, c.mask, datt, dq, dk, dv, batch=B, heads=heads, sq=Sq, sk=Sk, head_dim=d,
                     scale=1.0 / math.sqrt(d), dropout_p=rt.p, seed=c.seed_att)
        if c.mode == "self":
            rt.wgrad(ops.colsum, dqkv, g.b)                     # one pass over the [rows, 3H] gradient buffer
        else:
            rt.wgrad(ops.colsum, dq, g.rows_b(0, 1))
            if c.mode == "cross" and hasattr(c, "v_bf"):
                rt.wgrad(ops.colsum, dk, g.rows_b(1, 2))
                rt.wgrad(ops.colsum, dv, g.rows_b(2, 3))
            elif c.mode == "cross":
                rt.wgrad(ops.colsum, dkvb, g.rows_b(1, 3))
    dx = None
    if c.mode == "self":
        rt.wgrad_gemm(dqkv, c.x_bf, g.w)
        if need_dx:
            dx = _resid_gemm(dqkv, lp.w, M, H, 3 * H, dev, rt=rt, b_layout=1, resid=ds_f32)
    else:
        rt.wgrad_gemm(dq, c.x_bf, g.rows_w(0, 1))
        if need_dx:
            dx = _empty(M, H, _F32, dev)
            ops.gemm(dq, lp.rows(0, 1)[0], b_layout=1, resid=ds_f32, out_f32=dx)
        if c.mode == "cross":
            Mk = B * Sk
            if hasattr(c, "v_bf"):
                rt.wgrad_gemm(dk, c.kv_bf, g.rows_w(1, 2))
                rt.wgrad_gemm(dv, c.v_bf, g.rows_w(2, 3))
                dkv_src = _empty(Mk, H, _F32, dev)
                dv_src = _empty(Mk, H, _F32, dev)
                ops.gemm(dk, lp.rows(1, 2)[0], b_layout=1, out_f32=dkv_src)
                ops.gemm(dv, lp.rows(2, 3)[0], b_layout=1, out_f32=dv_src)
            else:
                rt.wgrad_gemm(dkvb, c.kv_bf, g.rows_w(1, 3))
                dkv_src = _empty(Mk, H, _F32, dev)
                ops.gemm(dkvb, lp.rows(1, 3)[0], b_layout=1, out_f32=dkv_src)
    # (in "kv" mode the K/V weights belong to the caller's batched projection)
    for (lin, (gw, gb)) in zip((mh.linear_q, mh.linear_k, mh.linear_v), g.per_param()):
        grads[lin.weight], grads[lin.bias] = gw, gb
    return dx, dkv_src, dv_src, grads


# ------------------------------------------------------------------------------------------
# MLP / FFN sub-layer:  LN(x + dropout(W2 dropout(relu(W1 x))))   (mca.py:123-125, net_utils.py:25-45)
# With norm=None: the bare MLP module (fp32 output, no residual).
# ------------------------------------------------------------------------------------------
def mlp_fwd(rt, mlp, x, norm=None):
    dev = x.bf.device
    M = x.bf.shape[0]
    c = Bag()
    force = _force(rt.p > 0 or rt.grad)
    lp1 = mlp.lp_fc().get(force, rt.split)
    lp2 = mlp.lp_out().get(force, rt.split)
    c.lp1, c.lp2, c.x_bf = lp1, lp2, x.bf
    p_mid = rt.p if mlp.fc.dropout_r > 0 else 0.0
    c.p_mid = p_mid
    hmid = _empty(M, lp1.n, _BF16, dev)
    hlo = _empty(M, lp1.n, _BF16, dev) if rt.split else None
    c.seed_mid = rt.seed()
    _mm(rt, x, lp1, 0, 1, relu=mlp.fc.use_relu, dropout_p=p_mid, seed=c.seed_mid, out_bf16=hmid, out_lo=hlo)
    c.hmid = hmid
    ha = Act(None, hmid, hlo)
    if norm is None:
        if lp2.n % 8 == 0:
            out = _empty(M, lp2.n, _F32, dev)
        else:   # narrow head (e.g. AttFlat glimpses): padded fp32 output buffer
            ldp = (lp2.n + 3) // 4 * 4
            out = torch.empty((M, ldp), dtype=_F32, device=dev)[:, :lp2.n]
        _mm(rt, ha, lp2, 0, 1, out_f32=out)
        return out, c
    # split-K (see _resid_gemm) only when training: fp32 atomics make the result depend on the
    # arrival order (1e-7 relative), and inference must stay bit-reproducible run to run
    use_sk = SPLITK_MIN_K > 0 and M <= SPLITK_MAX_ROWS and lp1.n >= SPLITK_MIN_K and rt.grad and SPLITK_FWD
    c.seed_out = rt.seed()
    if not use_sk and _can_fuse_ln(rt, M, lp2.n, lp2):
        out, c.s, c.mean, c.sigma = gemm_ln_fwd(hmid, lp2, norm, x.f32, rt.p, c.seed_out)
        return out, c
    s = torch.zeros((M, lp2.n), dtype=_F32, device=dev) if use_sk else _empty(M, lp2.n, _F32, dev)
    _mm(rt, ha, lp2, 0, 1, dropout_p=rt.p, seed=c.seed_out, resid=x.f32, out_f32=s, accumulate=use_sk)
    out, c.mean, c.sigma = ln_fwd(norm, s, split=rt.split)
    c.s = s
    return out, c


def mlp_bwd(rt, mlp, c, dout, norm=None, need_dx=True):
    """Returns (dx_f32, grads)."""
    dev = dout.device
    M = c.x_bf.shape[0]
    grads = {}
    g1, g2 = GradBuf(rt, c.lp1), GradBuf(rt, c.lp2)
    if norm is None:
        ds_f32 = None
        ds_bf = _bf_padded(M, c.lp2.n, dev)
        if ds_bf.stride(0) == c.lp2.n:
            ops.cast_bf16(dout.contiguous(), ds_bf)
        else:
            ds_bf.copy_(dout)
        rt.wgrad(ops.colsum, dout if dout.stride(-1) == 1 and dout.stride(0) % 4 == 0 else dout.contiguous(), g2.b)
    else:
        ds_f32, ds_bf, da2, db2 = ln_bwd(rt, norm, dout, c.s, c.mean, c.sigma, p=rt.p, seed=c.seed_out, dbias=g2.b)
        grads[norm.a_2], grads[norm.b_2] = da2, db2
    rt.wgrad_gemm(ds_bf, c.hmid, g2.w)
    dh = _empty(M, c.lp1.n, _BF16, dev)
    gate = c.hmid if (mlp.fc.use_relu or c.p_mid > 0) else None
    gate_scale = 1.0 / (1.0 - c.p_mid) if c.p_mid > 0 else 1.0
    if gate is not None and not mlp.fc.use_relu:
        raise ops.capi.McanError("dropout without ReLU in FC is not supported by the fused gate")
    # the FFN1 bias gradient (column sums of dh) comes out of the same epilogue, from the fp32 values
    ops.gemm(ds_bf, c.lp2.w, b_layout=1, gate=gate, gate_scale=gate_scale, out_bf16=dh, colsum=g1.b)
    rt.wgrad_gemm(dh, c.x_bf, g1.w)
    dx = None
    if need_dx:
        dx = _resid_gemm(dh, c.lp1.w, M, c.lp1.k, c.lp1.n, dev, rt=rt, b_layout=1, resid=ds_f32)
    (w1, b1), = g1.per_param()
    (w2, b2), = g2.per_param()
    grads[mlp.fc.linear.weight], grads[mlp.fc.linear.bias] = w1, b1
    grads[mlp.linear.weight], grads[mlp.linear.bias] = w2, b2
    return dx, grads


# ------------------------------------------------------------------------------------------
# SA / SGA layers
# ------------------------------------------------------------------------------------------
def sa_fwd(rt, sa, x, B, S, mask):
    y, c1 = att_fwd(rt, sa.mhatt, x, B, S, key_mask=mask, norm=sa.norm1)
    z, c2 = mlp_fwd(rt, sa.ffn.mlp, y, norm=sa.norm2)
    return z, (c1, c2)


def sa_bwd(rt, sa, ctx, dz):
    c1, c2 = ctx
    rt.begin_group()
    dy, grads = mlp_bwd(rt, sa.ffn.mlp, c2, dz, norm=sa.norm2)
    dx, _, _, g1 = att_bwd(rt, sa.mhatt, c1, dy, norm=sa.norm1)
    rt.end_group()
    grads.update(g1)
    return dx, grads


def sga_fwd(rt, sga, x, y, B, Sx, Sy, x_mask, y_mask, kv=None):
    """x: image-side Act [B*Sx,H]; y: question-side Act [B*Sy,H] (or kv=(K,V) pre-projected)."""
    a, c1 = att_fwd(rt, sga.mhatt1, x, B, Sx, key_mask=x_mask, norm=sga.norm1)
    if kv is None:
        b, c2 = att_fwd(rt, sga.mhatt2, a, B, Sx, kv_src=y, Sk=Sy, key_mask=y_mask, norm=sga.norm2)
    else:
        b, c2 = att_fwd(rt, sga.mhatt2, a, B, Sx, Sk=Sy, key_mask=y_mask, kv=kv, norm=sga.norm2)
    z, c3 = mlp_fwd(rt, sga.ffn.mlp, b, norm=sga.norm3)
    return z, (c1, c2, c3)


def sga_bwd(rt, sga, ctx, dz, dkv=None):
    c1, c2, c3 = ctx
    rt.begin_group()
    db, grads = mlp_bwd(rt, sga.ffn.mlp, c3, dz, norm=sga.norm3)
    da, dy, _, g2 = att_bwd(rt, sga.mhatt2, c2, db, norm=sga.norm2, dkv=dkv)
    dx, _, _, g1 = att_bwd(rt, sga.mhatt1, c1, da, norm=sga.norm1)
    rt.end_group()
    grads.update(g2)
    grads.update(g1)
    return dx, dy, grads


# ------------------------------------------------------------------------------------------
# MCA_ED (mca.py:178-186) as ONE kernel chain with the cross-layer K/V batch
# ------------------------------------------------------------------------------------------
def mca_ed_fwd(rt, m, x32, y32, B, Sx, Sy, x_mask, y_mask):
    """x32: question features fp32 [B*Sx, H]; y32: image features fp32 [B*Sy, H]."""
    H = m.hidden_size
    L = len(m.dec_list)
    dev = x32.device
    with refresh_scope(m.all_lps(), rt.p > 0 or rt.grad):
        return _mca_ed_fwd(rt, m, x32, y32, B, Sx, Sy, x_mask, y_mask, H, L, dev)


def _mca_ed_fwd(rt, m, x32, y32, B, Sx, Sy, x_mask, y_mask, H, L, dev):
    x = act_from_f32(x32, rt.split)
    y = act_from_f32(y32, rt.split)
    enc_ctx = []
    for enc in m.enc_list:
        x, c = sa_fwd(rt, enc, x, B, Sx, x_mask)
        enc_ctx.append(c)
    # K/V of the final encoder output for all decoder layers: one GEMM [B*Sx, H] x [H, L*2H]
    lpkv = m.lp_kv_all().get(_force(rt.p > 0 or rt.grad), rt.split)
    kv_all = _empty(B * Sx, 2 * H * L, _BF16, dev)
    kv_lo = _empty(B * Sx, 2 * H * L, _BF16, dev) if rt.split else None
    if L > 0:
        _mm(rt, x, lpkv, 0, len(lpkv.sizes), out_bf16=kv_all, out_lo=kv_lo)
    dec_ctx = []
    for i, dec in enumerate(m.dec_list):
        kv = (kv_all[:, 2 * H * i: 2 * H * i + H], kv_all[:, 2 * H * i + H: 2 * H * (i + 1)])
        if rt.split:
            kv = kv + (kv_lo[:, 2 * H * i: 2 * H * i + H], kv_lo[:, 2 * H * i + H: 2 * H * (i + 1)])
        y, c = sga_fwd(rt, dec, y, x, B, Sy, Sx, y_mask, x_mask, kv=kv)
        dec_ctx.append(c)
    ctx = Bag()
    ctx.enc, ctx.dec, ctx.lpkv, ctx.xenc_bf, ctx.B, ctx.Sx, ctx.Sy = enc_ctx, dec_ctx, lpkv, x.bf, B, Sx, Sy
    return x.f32, y.f32, ctx


# Deferred weight gradients (experiment, off by default).  The decoder's wgrad GEMMs (a third of its
# GEMM work) feed nothing in the backward chain, while the encoder backward that follows is a chain
# of ~110 small kernels on 896 rows that leaves most SMs idle.  With OVERLAP_WGRAD the decoder
# wgrads are queued and run on a second stream NEXT TO the encoder backward, on at most WGRAD_SMS
# SMs with the dynamic tile schedule (the remaining SMs stay free for the encoder chain).
# Measured on B200 (MCAN-large, batch 64): 9.33 ms/step with, 9.34 ms without -- both phases are
# bound by L2 -> shared-memory operand traffic, not by SM count, so sharing the GPU buys nothing.
OVERLAP_WGRAD = os.environ.get("MCAN_OVERLAP_WGRAD", "0") != "0"
WGRAD_SMS = int(os.environ.get("MCAN_WGRAD_SMS", "108"))
# Data parallel: hand the decoder layers' gradients to the all-reduce only once the whole decoder
# backward is enqueued (one large exchange next to the latency-bound encoder backward) instead of
# layer by layer next to the decoder's large GEMMs.
DP_DELAY_DECODER = os.environ.get("MCAN_DP_DELAY_DECODER", "0") != "0"
_side_streams = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=key)
    return st


def module_params(m):
    """list(m.parameters()), cached on the module: nn.Module.parameters() walks the whole module tree in Python
    (2.5 ms for MCA_ED, paid several times per eager training step).  The cache holds the Parameter objects, which
    survive .cuda() / .to() / load_state_dict(); a module whose parameters are REPLACED after the first forward
    must call invalidate_module_params."""
    cache = m.__dict__.get("_mcan_params")
    if cache is None:
        cache = list(m.parameters())
        m.__dict__["_mcan_params"] = cache
    return cache


def invalidate_module_params(root):
    for sub in root.modules():
        sub.__dict__.pop("_mcan_params", None)


def mca_ed_bwd(rt, m, ctx, dx_out, dy_out, after_layer=None):
    """dx_out/dy_out: fp32 grads of the two outputs.  after_layer(bufs, grads, kind) is called with the flat
    gradient buffers and the {parameter: gradient} map of a layer ("dec" | "kv" | "enc") as soon as its
    kernels are enqueued (hook for the overlapped gradient all-reduce in dp.py, or for the overlapped
    optimiser step on a single GPU, optim.EarlyStep)."""
    H = m.hidden_size
    L = len(m.dec_list)
    B, Sx, Sy = ctx.B, ctx.Sx, ctx.Sy
    dev = dy_out.device
    grads = {}
    params = module_params(m)
    total = sum(p.numel() for p in params) + 8 * len(params) + 64
    # weight gradients of the SA / SGA layers (all 2-D parameters except the K/V projections batched across the
    # decoder layers) are written by grouped launches into an uninitialised arena; everything else starts from zero
    batched = set(id(w) for w, _ in ctx.lpkv.pairs) if L > 0 else set()
    store = sum(p.numel() + 8 for p in params if p.dim() == 2 and id(p) not in batched) if GROUP_WGRADS else 0
    # split-K outputs of the question-side chain (2 per encoder layer + the batched K/V back-projection): one zero pool
    rows_x = B * Sx
    scratch = (2 * len(m.enc_list) + 1) * ((rows_x * H + 3) // 4 * 4) if (SPLITK_MIN_K > 0 and rows_x <= SPLITK_MAX_ROWS) else 0
    rt.use_arena(total - store if store and STORE_WGRADS else total, dev, store_numel=store, scratch_numel=scratch)
    dkv_all = _empty(B * Sx, 2 * H * L, _BF16, dev)
    dy = dy_out
    overlap = OVERLAP_WGRAD and L > 0 and len(m.enc_list) > 0
    pending = []        # per decoder layer: (deferred wgrad work, flat gradient buffers)
    for i in range(L - 1, -1, -1):
        dec = m.dec_list[i]
        dkv = (dkv_all[:, 2 * H * i: 2 * H * i + H], dkv_all[:, 2 * H * i + H: 2 * H * (i + 1)])
        if overlap:
            rt.deferred = []
        dy, _, g = sga_bwd(rt, dec, ctx.dec[i], dy, dkv=dkv)
        grads.update(g)
        if overlap:
            pending.append((rt.deferred, rt.drain()))
            rt.deferred = None
        elif after_layer is not None and not DP_DELAY_DECODER:
            after_layer(rt.drain(), g, "dec")
    dx = dx_out
    if L > 0:
        gkv = GradBuf(rt, ctx.lpkv)
        if overlap:
            rt.deferred = []
        rt.wgrad(ops.colsum, dkv_all, gkv.b)
        rt.wgrad_gemm(dkv_all, ctx.xenc_bf, gkv.w)
        dx = _resid_gemm(dkv_all, ctx.lpkv.w, B * Sx, H, 2 * H * L, dev, rt=rt, b_layout=1, resid=dx_out)
        gk = {}
        for (w, b), (gw, gb) in zip(ctx.lpkv.pairs, gkv.per_param()):
            gk[w], gk[b] = gw, gb
        grads.update(gk)
        if overlap:
            pending.append((rt.deferred, rt.drain()))
            rt.deferred = None
        elif after_layer is not None:
            after_layer(rt.drain(), gk, "kv")
    side = None
    if overlap:
        # fork: everything the deferred work reads has been produced on the current stream
        main = torch.cuda.current_stream()
        side = _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side), ops.launch_config(sm_limit=WGRAD_SMS, dynamic=True):
            for work, bufs in pending:
                for fn, args, kw in work:
                    fn(*args, **kw)
                if after_layer is not None:
                    after_layer(bufs)       # the all-reduce is ordered after this stream's work
    # while the deferred wgrads run, the encoder chain sizes its (persistent) grids for the SMs they leave free
    enc_limit = max(ops.num_sms_physical() - WGRAD_SMS, 16) if side is not None else None
    with ops.launch_config(sm_limit=enc_limit):
        for i in range(len(m.enc_list) - 1, -1, -1):
            dx, g = sa_bwd(rt, m.enc_list[i], ctx.enc[i], dx)
            grads.update(g)
            if after_layer is not None:
                after_layer(rt.drain(), g, "enc" if i > 0 else "enc_last")
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)    # join; `pending` kept every operand alive until here
        del pending
    return dx, dy, grads


# ------------------------------------------------------------------------------------------
# AttFlat (net.py:38-55)
# ------------------------------------------------------------------------------------------
def attflat_fwd(rt, af, x, B, S, mask):
    """x: Act [B*S, H] -> (x_atted fp32 [B,O], att_w fp32 [B,S,G], ctx)."""
    H, G, M = af.hidden_size, af.flat_glimpses, af.flat_mlp_size
    dev = x.bf.device
    c = Bag()
    force = _force(rt.p > 0 or rt.grad)
    lp1 = af.mlp.lp_fc().get(force, rt.split)
    lpm = af.lp_merge().get(force, rt.split)
    c.lp1, c.lpm, c.x, c.B, c.S, c.mask = lp1, lpm, x, B, S, mask
    p_mid = rt.p if af.mlp.fc.dropout_r > 0 else 0.0
    c.p_mid = p_mid
    hmid = _empty(B * S, M, _BF16, dev)
    hlo = _empty(B * S, M, _BF16, dev) if rt.split else None
    c.seed_mid = rt.seed()
    _mm(rt, x, lp1, 0, 1, relu=True, dropout_p=p_mid, seed=c.seed_mid, out_bf16=hmid, out_lo=hlo)
    att_w = torch.empty((B, S, G), dtype=_F32, device=dev)
    pooled = _empty(B, G * H, _BF16, dev)
    w2 = af.mlp.linear.weight.detach()
    b2 = af.mlp.linear.bias.detach()
    out = _empty(B, lpm.n, _F32, dev)
    p32 = _empty(B, G * H, _F32, dev)          # fp32 pooled sums: the backward's softmax term is dpooled . pooled
    if rt.split:
        ops.attflat_pool_fwd(hmid, w2, b2, mask, x.f32, batch=B, s=S, h=H, mlp=M, glimpses=G, att_w=att_w,
                             pooled_f32=p32, hmid_lo=hlo)
        _mm(rt, act_from_f32(p32, True), lpm, 0, 1, out_f32=out)
    else:
        ops.attflat_pool_fwd(hmid, w2, b2, mask, x.f32, batch=B, s=S, h=H, mlp=M, glimpses=G, att_w=att_w,
                             pooled_f32=p32, pooled_bf16=pooled)
        ops.gemm(pooled, lpm.w, bias=lpm.b, out_f32=out)
    c.hmid, c.att_w, c.pooled, c.pooled32 = hmid, att_w, pooled, p32
    return out, att_w, c


def attflat_bwd(rt, af, c, dout, need_dx=True):
    H, G, M = af.hidden_size, af.flat_glimpses, af.flat_mlp_size
    dev = dout.device
    B, S = c.B, c.S
    grads = {}
    gm, g1 = GradBuf(rt, c.lpm), GradBuf(rt, c.lp1)
    dout = dout.contiguous()
    dout_bf = _empty(B, c.lpm.n, _BF16, dev)
    ops.cast_bf16(dout, dout_bf)
    ops.colsum(dout, gm.b)
    ops.gemm(dout_bf, c.pooled, a_layout=1, b_layout=1, out_f32=gm.w, accumulate=True)
    dpooled = _empty(B, G * H, _F32, dev)
    ops.gemm(dout_bf, c.lpm.w, b_layout=1, out_f32=dpooled)
    dx = _empty(B * S, H, _F32, dev)
    dh = _empty(B * S, M, _BF16, dev)
    gw2 = rt.zeros(G * M + G, dev)
    ops.attflat_pool_bwd(dpooled, c.pooled32, c.hmid, af.mlp.linear.weight.detach(), c.mask, c.x.f32, c.att_w, batch=B,
                         s=S, h=H, mlp=M, glimpses=G, gate_scale=1.0 / (1.0 - c.p_mid) if c.p_mid > 0 else 1.0,
                         dx=dx, dhmid=dh, dw2=gw2[: G * M], db2=gw2[G * M:])
    ops.colsum(dh, g1.b)
    ops.gemm(dh, c.x.bf, a_layout=1, b_layout=1, out_f32=g1.w, accumulate=True)
    if need_dx:
        ops.gemm(dh, c.lp1.w, b_layout=1, resid=dx, out_f32=dx)   # in place: dx += dh W1
    (w1, b1), = g1.per_param()
    (wm, bm), = gm.per_param()
    grads[af.mlp.fc.linear.weight], grads[af.mlp.fc.linear.bias] = w1, b1
    grads[af.mlp.linear.weight], grads[af.mlp.linear.bias] = gw2[: G * M].view(G, M), gw2[G * M:]
    grads[af.linear_merge.weight], grads[af.linear_merge.bias] = wm, bm
    return (dx if need_dx else None), grads


# ------------------------------------------------------------------------------------------
# plain linear on the tcgen05 GEMM (img_feat_linear / proj: the rows next to the hot path)
# ------------------------------------------------------------------------------------------
def linear_fwd(lp, x32, split=False, mask_out=None):
    """mask_out (uint8 [rows], optional): also receives the reference's zero-row mask of x32 (net.py:135-137),
    computed in the same pass that casts x32 to the bf16 GEMM operand."""
    dev = x32.device
    M = x32.shape[0]
    c = Bag()
    xbf = _bf_padded(M, lp.k, dev)
    xlo = _bf_padded(M, lp.k, dev) if split else None
    if mask_out is not None:
        ops.rowmask_cast(x32.contiguous(), xbf, xlo, mask_out)
    elif xbf.stride(0) == lp.k:
        ops.cast_bf16(x32.contiguous(), xbf, xlo)
    else:
        ops.rowmask_cast(x32.contiguous(), xbf, xlo)
    ldo = (lp.n + 3) // 4 * 4
    out = torch.empty((M, ldo), dtype=_F32, device=dev)[:, :lp.n]
    if split:
        ops.gemm([xbf, xbf, xlo], [lp.w, lp.w_lo, lp.w], bias=lp.b, out_f32=out)
    else:
        ops.gemm(xbf, lp.w, bias=lp.b, out_f32=out)
    c.xbf, c.lp = xbf, lp
    return out, c


def linear_bwd(rt, c, dout, need_dx=True):
    lp = c.lp
    dev = dout.device
    M = dout.shape[0]
    g = GradBuf(rt, lp)
    dbf = _bf_padded(M, lp.n, dev)
    if dbf.stride(0) == lp.n and dout.is_contiguous():
        ops.cast_bf16(dout, dbf)
    else:
        dbf.copy_(dout)
    ops.colsum(dbf, g.b)
    ops.gemm(dbf, c.xbf, a_layout=1, b_layout=1, out_f32=g.w, accumulate=True)
    dx = None
    if need_dx:
        dx = _empty(M, lp.k, _F32, dev)
        ops.gemm(dbf, lp.w, b_layout=1, out_f32=dx)
    (gw, gb), = g.per_param()
    return dx, gw, gb


# ------------------------------------------------------------------------------------------
# output head: proj_norm(lang [+ img]) -> proj -> sigmoid [-> BCELoss(sum)]
# (net.py:125-129, 182-184; exec.py:67,178)
# ------------------------------------------------------------------------------------------
def head_fwd(rt, norm, lp, x, x2, target):
    """x, x2 (or None): fp32 [B, O]; target fp32 [B, A] or None.
    Returns (a fp32 [B,O], probs fp32 [B,A], loss 0-dim or None, ctx)."""
    dev = x.device
    B, O = x.shape
    c = Bag()
    y32 = _empty(B, O, _F32, dev)
    ybf = _empty(B, O, _BF16, dev)
    ylo = _empty(B, O, _BF16, dev) if rt.split else None
    c.mean = torch.empty(B, dtype=_F32, device=dev)
    c.sigma = torch.empty(B, dtype=_F32, device=dev)
    if x2 is None:
        c.s = x
        ops.layernorm_fwd(x, norm.a_2.detach(), norm.b_2.detach(), norm.eps, y_f32=y32, y_bf16=ybf, y_lo=ylo,
                          mean=c.mean, sigma=c.sigma)
    else:
        c.s = _empty(B, O, _F32, dev)
        ops.layernorm_add_fwd(x, x2, norm.a_2.detach(), norm.b_2.detach(), norm.eps, s_out=c.s, y_f32=y32,
                              y_bf16=ybf, y_lo=ylo, mean=c.mean, sigma=c.sigma)
    ldo = (lp.n + 3) // 4 * 4
    if rt.split:
        logits = torch.empty((B, ldo), dtype=_F32, device=dev)[:, :lp.n]
        ops.gemm([ybf, ybf, ylo], [lp.w, lp.w_lo, lp.w], bias=lp.b, out_f32=logits)
    elif rt.grad and (SPLITK_FWD or SPLITK_HEAD) and SPLITK_MIN_K > 0 and B <= SPLITK_MAX_ROWS and lp.k >= SPLITK_MIN_K:
        # 64 rows x 3129 answers x K 2048 is one wave of 49 tiles that stream the 12.8 MB weight through 49 SMs
        # (45 us); split-K puts every SM on it (training only, like every split-K GEMM: fp32 atomics)
        logits = torch.zeros((B, ldo), dtype=_F32, device=dev)[:, :lp.n]
        ops.gemm(ybf, lp.w, bias=lp.b, out_f32=logits, accumulate=True)
    else:
        logits = torch.empty((B, ldo), dtype=_F32, device=dev)[:, :lp.n]
        ops.gemm(ybf, lp.w, bias=lp.b, out_f32=logits)
    probs = _empty(B, lp.n, _F32, dev)
    loss = torch.empty((), dtype=_F32, device=dev) if target is not None else None
    ops.sigmoid_bce_fwd(logits, probs, target, loss)
    c.abf, c.lp, c.probs, c.target = ybf, lp, probs, target
    return y32, probs, loss, c


def head_bwd(rt, norm, c, g_a, g_probs, g_loss):
    """Returns (ds fp32 [B,O] -- the gradient of both summands --, {parameter-role: gradient})."""
    lp = c.lp
    dev = c.probs.device
    B = c.probs.shape[0]
    g = GradBuf(rt, lp)
    dz = torch.empty((B, (lp.n + 7) // 8 * 8), dtype=_BF16, device=dev)[:, :lp.n]     # the kernel zeroes the pad columns
    if c.target is not None:
        ops.sigmoid_bce_bwd(c.probs, dz, target=c.target, gscale=g_loss, dbias=g.b)
    else:
        ops.sigmoid_bce_bwd(c.probs, dz, gout=g_probs.contiguous(), dbias=g.b)
    ops.gemm(dz, c.abf, a_layout=1, b_layout=1, out_f32=g.w, accumulate=True)
    da = _resid_gemm(dz, lp.w, B, lp.k, lp.n, dev, b_layout=1, resid=g_a)
    ds, _, da2, db2 = ln_bwd(rt, norm, da, c.s, c.mean, c.sigma, want_bf=False)
    (gw, gb), = g.per_param()
    return ds, gw, gb, da2, db2


# ------------------------------------------------------------------------------------------
# question encoder: embedding + LSTM (net.py:66-78, 99, 103-104) -- csrc/lstm.cu
# ------------------------------------------------------------------------------------------
LSTM_KERNEL = os.environ.get("MCAN_LSTM", "1") != "0"
LSTM_HIDDEN = (128, 256, 512, 1024)
LSTM_CHUNK = 64       # samples per persistent launch
LSTM_SPLIT_INPUT = os.environ.get("MCAN_LSTM_SPLIT_INPUT", "1") != "0"


def lstm_supported(lstm, split):
    return (LSTM_KERNEL and not split and lstm.num_layers == 1 and not lstm.bidirectional and lstm.batch_first and
            lstm.bias and getattr(lstm, "proj_size", 0) == 0 and lstm.hidden_size in LSTM_HIDDEN and
            ops.num_sms_physical() >= 128)


def qenc_fwd(rt, table, lp_ih, lp_hh, tokens, training):
    """tokens int64 [B, T] -> (q fp32 [B*T, H], mask uint8 [B*T] (token == 0), ctx)."""
    dev = tokens.device
    B, T = tokens.shape
    S1 = T + 1
    H, E = lp_hh.k, lp_ih.k
    R = B * S1
    c = Bag()
    ldx = lp_ih.w.stride(0)         # E rounded up to a multiple of 64: no ragged 64-column chunk in any GEMM of the encoder
    x = torch.empty((R, ldx), dtype=_BF16, device=dev)
    mask = torch.empty(B * T, dtype=torch.uint8, device=dev)
    xw = _empty(R, 4 * H, _F32, dev)
    if LSTM_SPLIT_INPUT and lp_ih.w_lo is not None:
        # The input projection feeds all T steps of the recurrence: it runs at split precision (x and W_ih as bf16
        # hi + lo, three tensor-core products) -- K = 300, a few microseconds -- so that the only bf16 rounding inside
        # the question encoder is that of W_hh and h in the recurrent product.
        xlo = torch.empty((R, ldx), dtype=_BF16, device=dev)
        ops.embed_gather(tokens, table, x, mask, xlo)
        ops.gemm([x[:, :E], x[:, :E], xlo[:, :E]], [lp_ih.w, lp_ih.w_lo, lp_ih.w], bias=lp_ih.b, out_f32=xw)
    else:
        ops.embed_gather(tokens, table, x, mask)
        ops.gemm(x[:, :E], lp_ih.w, bias=lp_ih.b, out_f32=xw)
    hbuf = _empty(R, H, _BF16, dev)
    q = _empty(B * T, H, _F32, dev)
    cbuf = _empty(R, H, _F32, dev) if training else None
    gates = _empty(R, 4 * H, _F32, dev) if training else None
    for b0 in range(0, B, LSTM_CHUNK):
        nb = min(LSTM_CHUNK, B - b0)
        r0, r1 = b0 * S1, (b0 + nb) * S1
        ops.lstm_fwd(xw[r0:r1], lp_hh.w, lp_hh.b, hbuf[r0:r1], q[b0 * T:(b0 + nb) * T],
                     cbuf[r0:r1] if training else None, gates[r0:r1] if training else None, batch=nb, steps=T, hidden=H)
    c.tokens, c.x, c.hbuf, c.cbuf, c.gates, c.lp_ih, c.lp_hh, c.B, c.T, c.E = tokens, x, hbuf, cbuf, gates, lp_ih, lp_hh, B, T, E
    return q, mask, c


def qenc_bwd(rt, c, dq, vocab):
    """dq fp32 [B*T, H] -> (dTable [V, E], dW_ih, db_ih, dW_hh, db_hh)."""
    dev = dq.device
    B, T, E = c.B, c.T, c.E
    S1 = T + 1
    H = c.lp_hh.k
    R = B * S1
    da = _empty(R, 4 * H, _BF16, dev)
    for b0 in range(0, B, LSTM_CHUNK):
        nb = min(LSTM_CHUNK, B - b0)
        r0, r1 = b0 * S1, (b0 + nb) * S1
        ops.lstm_bwd(dq[b0 * T:(b0 + nb) * T], c.lp_hh.w, c.hbuf[r0:r1], c.cbuf[r0:r1], c.gates[r0:r1], da[r0:r1],
                     batch=nb, steps=T, hidden=H)
    g_ih, g_hh = GradBuf(rt, c.lp_ih), GradBuf(rt, c.lp_hh)
    # dW_hh = dA^T h_prev and dW_ih = dA^T x as one grouped launch.  E = 300 is not a multiple of the epilogue's 64-column
    # chunk: the GEMMs run on the zero-padded width (x and the W_ih copy are E_pad = 320 wide), so no chunk takes the
    # per-element path (it cost 25 us per launch); dW_ih lands in a padded temporary and is copied out.
    ldx = c.x.shape[1]
    dwi = torch.empty((4 * H, ldx), dtype=_F32, device=dev)
    ops.gemm_grouped([(da, c.hbuf, g_hh.w), (da, c.x, dwi)], accumulate=False)
    g_ih.w.copy_(dwi[:, :E])
    ops.colsum(da, g_hh.b)
    ops.colsum(da, g_ih.b)
    dx = _resid_gemm(da, c.lp_ih.w_full, R, ldx, 4 * H, dev, b_layout=1)[:, :E]      # short M, K = 4H: split-K on all SMs
    dtable = rt.zeros(vocab * E, dev).view(vocab, E)
    ops.embed_scatter_add(c.tokens, dx, dtable)
    return dtable, g_ih.w, g_ih.b, g_hh.w, g_hh.b
