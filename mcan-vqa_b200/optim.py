"""Fused multi-tensor AdamW on the library's `mcan_adamw_multi` kernel (SURVEY 8f "next" #3).

Semantics = torch.optim.AdamW as the reference builds it (core/model/optim.py:58-64: lr set per
step by WarmupOptimizer, betas (0.9, 0.999), eps 1e-8, weight_decay 1e-4): decoupled weight decay,
bias-corrected moments.  ONE kernel per step updates every parameter and re-emits the bf16
GEMM-operand copies ("shadows") that the next forward reads, so the separate fp32 -> bf16 refresh
pass over all weights disappears.  Learning rate and step count are device scalars: a captured
CUDA graph replays the step with new values.

Differences from torch.optim.AdamW, all deliberate: one global step counter (parameters that never
receive a gradient are skipped, as in torch, but a parameter that only sometimes receives one
shares the counter); no amsgrad / maximize / foreach options.  `state_dict()` uses torch's layout,
so checkpoints written by core/exec.py:237-244 load into either optimiser.
"""
import torch

from . import capi, ops

_F32 = torch.float32
_CHUNK = 4096
_F32_FLAG = 1 << 62
_F32_FLAG2 = 1 << 61
_BF16_GRAD = 1 << 60
_WORDS = 8


class FusedAdamW(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamW: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise capi.McanError("FusedAdamW needs CUDA parameters (the MCAN hot path has no CPU fallback)")
        for p in self.params:
            if p.device != dev or p.dtype != _F32 or not p.is_contiguous():
                raise capi.McanError("FusedAdamW: parameters must be contiguous fp32 tensors on one device")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=self.params)]
        self.device = dev
        self.lr_t = torch.zeros((), dtype=_F32, device=dev)
        self.step_t = torch.zeros((), dtype=_F32, device=dev)
        # flat moment buffers; every parameter's slice starts 16-byte aligned
        self._off, tot = [], 0
        for p in self.params:
            self._off.append(tot)
            tot += (p.numel() + 3) // 4 * 4
        self._m = torch.zeros(tot, dtype=_F32, device=dev)
        self._v = torch.zeros(tot, dtype=_F32, device=dev)
        self._shadow = {}       # id(param) -> up to two contiguous bf16 / fp32 tensors of the same numel
        self._managed = []      # LinearParams whose operand copies this optimiser keeps current
        self._tables = {}
        self._captured = []
        # pinned staging for the segment tables, allocated up front: a table built while a CUDA graph is
        # being captured must not allocate host memory, and the captured upload re-reads its buffer
        # on every replay, so buffers are never recycled
        self._pinned = [torch.empty((len(self.params), _WORDS), dtype=torch.int64).pin_memory() for _ in range(8)]
        self.epoch = 0          # bumped on every step (and by note_replay): low-order copies go stale
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._touched = set()   # indices of the parameters that have been updated at least once (state_dict emits only these)
        self._early_done = None  # ids of the parameters already updated in this step by an EarlyStep (None: no early step)

    # -- operand copies ------------------------------------------------------------------------
    def attach_shadows(self, lps):
        """Keep the bf16 operand copies of these blocks.LinearParams in sync with the masters."""
        by_id = {id(p): p for p in self.params}
        for lp in lps:
            if lp.managed is self:
                continue
            items = lp.shadow_items()
            if items is None or any(id(src) not in by_id for _, src in items):
                continue
            if any(len(self._shadow.get(id(src), ())) >= 2 for _, src in items):
                continue        # the kernel writes at most two copies per parameter
            for dst, src in items:
                self._shadow.setdefault(id(src), []).append(dst)
            lp.managed = self
            self._managed.append(lp)
        self._tables.clear()

    def note_replay(self):
        """A captured graph containing step() was replayed (no Python ran)."""
        self.epoch += 1

    # -- torch.optim.Optimizer surface used by core/exec.py and WarmupOptimizer ---------------------
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_()
                p.grad.zero_()

    def _table(self, active):
        key = tuple((i, self.params[i].data_ptr(), g.data_ptr()) for i, g in active)
        entry = self._tables.get(key)
        if entry is not None:
            return entry
        if len(self._tables) > 96:
            self._tables.clear()
        rows, chunk = [], 0
        for i, g in active:
            p = self.params[i]
            n, o = p.numel(), self._off[i]
            sh = self._shadow.get(id(p), [])
            flag = (_F32_FLAG if (len(sh) > 0 and sh[0].dtype == _F32) else 0) | \
                   (_F32_FLAG2 if (len(sh) > 1 and sh[1].dtype == _F32) else 0) | \
                   (_BF16_GRAD if g.dtype == torch.bfloat16 else 0)
            rows.append([p.data_ptr(), g.data_ptr(), self._m.data_ptr() + 4 * o, self._v.data_ptr() + 4 * o,
                         sh[0].data_ptr() if len(sh) > 0 else 0, n, chunk | flag,
                         sh[1].data_ptr() if len(sh) > 1 else 0])
            chunk += (n + _CHUNK - 1) // _CHUNK
        host = self._pinned.pop() if self._pinned else torch.empty((len(self.params), _WORDS), dtype=torch.int64).pin_memory()
        host = host[:len(rows)]
        host.copy_(torch.tensor(rows, dtype=torch.int64))
        table = torch.empty(host.shape, dtype=torch.int64, device=self.device)
        table.copy_(host, non_blocking=True)      # captured with the step when a graph is being recorded
        entry = (table, host, len(rows), chunk)
        self._tables[key] = entry
        if torch.cuda.is_current_stream_capturing():
            self._captured.append(entry)     # a graph replays the upload from `host`: keep it alive for good
        return entry

    def _prepare(self):
        """Learning rate and step count of this step into their device scalars (current stream)."""
        g0 = self.param_groups[0]
        lr = g0["lr"]
        if torch.is_tensor(lr):
            self.lr_t.copy_(lr.detach().reshape(()).to(_F32), non_blocking=True)
        else:
            self.lr_t.fill_(float(lr))
        if not torch.cuda.is_current_stream_capturing():
            while len(self._pinned) < 40:     # keep staging buffers ready for tables built during a capture
                self._pinned.append(torch.empty((len(self.params), _WORDS), dtype=torch.int64).pin_memory())
        self.step_t.add_(1.0)

    def _collect(self):
        active = []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or (self._early_done is not None and id(p) in self._early_done):
                continue
            if g.dtype != _F32 or not g.is_contiguous():
                g = g.to(_F32).contiguous()
            active.append((i, g))
        return active

    def _begin_step(self):
        if self._early_done is not None:     # an EarlyStep already prepared this step and updated some parameters
            return self._collect()
        active = self._collect()
        if active:
            self._prepare()
        return active

    def _launch(self, active, short_ctas=False):
        if not active:
            return
        g0 = self.param_groups[0]
        table, _host, nseg, chunks = self._table(active)
        self._touched.update(i for i, _ in active)
        b1, b2 = g0["betas"]
        lib = capi.load()
        capi.check(lib.mcan_adamw_multi(table.data_ptr(), nseg, chunks, self.lr_t.data_ptr(), self.step_t.data_ptr(),
                                        float(b1), float(b2), float(g0["eps"]), float(g0["weight_decay"]),
                                        1 if short_ctas else 0, ops._stream()), "mcan_adamw_multi")

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise capi.McanError("FusedAdamW.step: closures are not supported")
        self._launch(self._begin_step())
        if self._early_done is not None:
            self._early_done = None
            _early.join()
        self.epoch += 1
        return None

    @torch.no_grad()
    def step_buckets(self, buckets):
        """Data-parallel step: `buckets` = [(work, tensors)] from dp.GradSync.take_buckets(), in the
        order the gradient all-reduces were launched.  Each bucket's parameters are updated as
        soon as ITS all-reduce has finished, so the update of the early buckets (the decoder layers)
        overlaps the all-reduce of the late ones (encoder, LSTM, embedding) instead of waiting behind
        them.  Every gradient must belong to some bucket."""
        remaining = self._begin_step()
        for entry in buckets:
            work, tensors = entry[0], entry[1]
            if len(entry) > 3 and entry[3]:
                continue      # already waited for and applied next to the encoder backward (EarlyStep.on_dp_layer)
            work.wait()       # the current stream waits for this all-reduce; no host sync
            view = getattr(work, "grad_view", None)      # bf16 exchange: the reduced values live in the exchange buffers
            pairs = entry[2] if len(entry) > 2 else None
            mine = []
            if pairs:
                # the bucket knows its (parameter, gradient) pairs -- including the bf16 weight gradients that never
                # become a .grad (blocks.WGRAD_BF16)
                taken = set()
                open_idx = set(i for i, _ in remaining)
                for prm, g in pairs:
                    i = self._index.get(id(prm))
                    if i is None or g is None or (self._early_done is not None and id(prm) in self._early_done):
                        continue
                    if g.dtype == _F32 and i not in open_idx:
                        continue      # no .grad (any more): already applied through another handle of this bucket
                    gv = view(g) if view is not None else g
                    if gv is None:
                        raise capi.McanError("FusedAdamW.step_buckets: a gradient of the bucket is not in its exchange buffers")
                    mine.append((i, gv))
                    taken.add(i)
                remaining = [(i, g) for i, g in remaining if i not in taken]
            else:
                spans = [(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in tensors]
                rest = []
                for i, g in remaining:
                    a = g.data_ptr()
                    if any(lo <= a < hi for lo, hi in spans):
                        mine.append((i, view(g) if view is not None else g))
                    else:
                        rest.append((i, g))
                remaining = rest
            self._launch(sorted(mine, key=lambda t: t[0]))
        if remaining:
            raise capi.McanError("FusedAdamW.step_buckets: %d gradients were not part of any all-reduce bucket"
                                 % len(remaining))
        if self._early_done is not None:
            self._early_done = None
            _early.join()
        self.epoch += 1
        return None

    # -- checkpoints (torch.optim layout) --------------------------------------------------------
    def _views(self, i):
        p, o = self.params[i], self._off[i]
        n = p.numel()
        return self._m[o:o + n].view_as(p), self._v[o:o + n].view_as(p)

    def state_dict(self):
        step = self.step_t.detach().clone().cpu()
        state = {}
        if float(step) > 0:
            # like torch.optim.AdamW: no state for a parameter that never received a gradient
            for i in sorted(self._touched):
                m, v = self._views(i)
                state[i] = {"step": step.clone(), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        g0 = self.param_groups[0]
        group = {k: v for k, v in g0.items() if k != "params"}
        group.update(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FusedAdamW.load_state_dict: parameter groups do not match")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        order = groups[0]["params"]
        steps = [0.0]
        self._touched = set()
        for pos, idx in enumerate(order):
            st = sd["state"].get(idx)
            m, v = self._views(pos)
            if st is None:
                m.zero_()
                v.zero_()
                continue
            m.copy_(st["exp_avg"])
            v.copy_(st["exp_avg_sq"])
            steps.append(float(st["step"]))
            self._touched.add(pos)
        if len(set(steps[1:])) > 1:
            import warnings
            warnings.warn("FusedAdamW.load_state_dict: per-parameter step counts differ (%s .. %s); this optimiser keeps ONE "
                          "step counter and continues from the largest" % (min(steps[1:]), max(steps[1:])))
        self.step_t.fill_(max(steps))


class EarlyStep(object):
    """Single-GPU overlap of the optimiser with the backward pass (the data-parallel path overlaps the
    gradient all-reduce there instead, dp.py).

    AdamW is pure HBM traffic (30 bytes per parameter, 0.95 ms for MCAN-large at 98 % of the HBM peak), and the
    encoder half of MCA_ED's backward is a chain of ~70 latency-bound kernels on 896 rows that leaves most SMs
    and nearly all of the HBM bandwidth idle.  MCA_ED.backward reports each finished layer (blocks.mca_ed_bwd
    `after_layer`); once the decoder layers and the batched K/V projection are done -- when the encoder chain
    starts -- their parameters are updated on a second, lower-priority stream with short-lived CTAs, then every
    encoder layer as it finishes; FusedAdamW.step() only updates what is left (LSTM, embedding, heads) and joins.
    The arithmetic is unchanged: every parameter sees exactly one AdamW update with the same gradient."""

    def __init__(self, opt):
        self.opt = opt
        self.side = torch.cuda.Stream(device=opt.device)        # default (= lowest) priority
        self.waiting = []

    def begin(self):
        """Before backward(): fixes this step's learning rate / step count (current stream)."""
        self.opt._prepare()
        self.opt._early_done = set()
        self.waiting = []
        self.dp_done = 0
        self.forked = False

    def on_layer(self, bufs, grads=None, kind="enc"):
        if grads is None or self.opt._early_done is None:
            return
        pairs = []
        for prm, g in grads.items():
            i = self.opt._index.get(id(prm))
            if i is None or g is None or prm.grad is not None:   # (an accumulated .grad: leave it to step())
                continue
            if g.dtype != _F32 or not g.is_contiguous():
                continue
            pairs.append((i, g.detach()))      # a fresh view: never keep autograd's own gradient tensor alive (see dp.on_bufs)
        self.waiting += pairs
        if kind == "dec":
            return          # the decoder backward saturates the GPU by itself: hold the update back
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self.opt._launch(sorted(self.waiting, key=lambda t: t[0]), short_ctas=True)
        for i, _ in self.waiting:
            self.opt._early_done.add(id(self.opt.params[i]))
        self.waiting = []

    def on_dp_layer(self, sync, kind):
        """Data parallel (called by dp.GradSync.on_bufs after every layer): once the encoder chain has started, every
        bucket whose all-reduce has been launched is applied on the side stream as soon as that all-reduce finishes
        -- the update of the decoder layers no longer waits for the end of the backward pass."""
        if kind == "dec" or self.opt._early_done is None:
            return
        todo = [k for k in range(self.dp_done, len(sync.pending)) if len(sync.pending[k]) == 3]
        if not todo:
            return
        if not self.forked:
            self.side.wait_stream(torch.cuda.current_stream())     # lr / step scalars of this step
            self.forked = True
        with torch.cuda.stream(self.side):
            for k in todo:
                work, tensors, pairs = sync.pending[k]
                active = []
                view = getattr(work, "grad_view", None)  # bf16 exchange: read the reduced values from the staging buffer
                for prm, g in pairs:
                    i = self.opt._index.get(id(prm))
                    if i is None or g is None or g.dtype not in (_F32, torch.bfloat16) or not g.is_contiguous() or \
                            (g.dtype != _F32 and view is None):
                        active = None
                        break
                    gv = view(g) if view is not None else g
                    if gv is None:
                        active = None
                        break
                    active.append((i, gv))
                if active is None:
                    continue          # left to step_buckets
                work.wait()           # the side stream waits for this all-reduce
                self.opt._launch(sorted(active, key=lambda t: t[0]), short_ctas=True)
                for i, _ in active:
                    self.opt._early_done.add(id(self.opt.params[i]))
                sync.pending[k] = (work, tensors, pairs, True)
        self.dp_done = len(sync.pending)

    def join(self):
        torch.cuda.current_stream().wait_stream(self.side)


_early = None


def set_early(es):
    global _early
    _early = es


def early_hook():
    """The per-layer callback MCA_ED's backward should use on a single GPU (None when inactive)."""
    if _early is not None and _early.opt._early_done is not None:
        return _early.on_layer
    return None
