"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce(SUM) over NVLink.

Replaces the reference's nn.DataParallel (core/exec.py:62-63): there, replicas are re-broadcast
every forward, outputs gathered on GPU 0 and gradients summed onto GPU 0
(reduce_add_coalesced).  Here every rank owns a persistent replica and its own batch shard;
the only exchange is ONE all-reduce(SUM) of the gradients per optimiser step -- SUM, not mean,
because the loss is BCELoss(reduction='sum') (core/exec.py:67), so the result equals the
single-process gradient of the global batch.

Two modes:
  * overlap (bench / training harness, grad_accu_steps == 1): MCA_ED's backward hands the flat
    gradient buffers of each finished layer to `on_bufs`, which launches a coalesced async
    all-reduce while the next layer's kernels run; parameters outside the backbone are caught
    by post-accumulate-grad hooks; an end-of-backward callback waits for everything.
  * at-step (reference exec.py unchanged, any grad_accu_steps): `sync_all_grads` reduces all
    .grad tensors once, called from the overlay WarmupOptimizer.step().
"""
import os

import torch
import torch.distributed as dist
from torch.autograd import Variable

_active = None


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _all_reduce_sum_async(tensors, group=None):
    """Async all-reduce(SUM) of a list of tensors; one coalesced NCCL launch on GPUs, one call per
    tensor on backends without coalescing (gloo, used by the CPU tests).  Returns waitable handles."""
    if dist.get_backend(group) == "nccl":
        with dist._coalescing_manager(group, device=tensors[0].device, async_ops=True) as cm:
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return [cm]
    return [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tensors]


class _CompressedWork(object):
    """All-reduce(SUM) of a bf16 copy of fp32 gradient buffers (MCAN_DP_COMPRESS=bf16, opt-in
    experiment): half the bytes on the wire (403 instead of 806 MB for MCAN-large; stand-alone on
    8 x B200 1.03 instead of 1.98 ms).  wait() makes the current stream wait for the exchange and
    writes the sums back into the fp32 buffers, so .grad, the optimiser and grad clipping see
    ordinary fp32 gradients.  The sum itself is rounded to bf16 at every reduction step -- not
    bit-faithful to the fp32 exchange, hence off by default."""

    def __init__(self, tensors, group):
        self.tensors = list(tensors)
        n = sum(t.numel() for t in self.tensors)
        self.stage = torch.empty(n, dtype=torch.bfloat16, device=self.tensors[0].device)
        o = 0
        for t in self.tensors:
            self.stage[o:o + t.numel()].copy_(t.view(-1))
            o += t.numel()
        self.work = dist.all_reduce(self.stage, op=dist.ReduceOp.SUM, group=group, async_op=True)
        self.done = False

    def wait(self):
        if self.done:
            return
        self.work.wait()
        o = 0
        for t in self.tensors:
            t.view(-1).copy_(self.stage[o:o + t.numel()])
            o += t.numel()
        self.done = True


class _Bf16Bucket(object):
    """bf16 exchange WITHOUT pack / unpack passes in PyTorch (MCAN_DP_COMPRESS=bf16 with the bucket-wise fused optimiser):
    the fp32 gradient buffers of a bucket are cast into a persistent bf16 staging buffer by the library's cast kernel
    (one launch per buffer, 2-3 per bucket), the staging buffer is all-reduced (half the bytes on NVLink and half the
    HBM traffic of the NCCL kernels next to the GEMMs), and FusedAdamW reads the reduced bf16 values straight from
    it (`grad_view`; segment-table bit 60) -- the sums are never written back, .grad keeps the rank-local fp32 values."""

    def __init__(self, tensors, group, stage):
        from . import ops
        self.spans = []
        o = 0
        direct = []       # bf16 buffers (weight gradients written by the grouped launches): reduced in place
        for t in tensors:
            if t.dtype == torch.bfloat16:
                direct.append(t)
                continue
            n = t.numel()
            ops.cast_bf16(t.view(-1), stage[o:o + n])
            self.spans.append((t.data_ptr(), t.data_ptr() + 4 * n, o))
            o += (n + 7) // 8 * 8
        self.stage = stage[:o]
        bufs = ([self.stage] if o > 0 else []) + direct
        self.works = _all_reduce_sum_async(bufs, group) if len(bufs) > 1 else \
            [dist.all_reduce(bufs[0], op=dist.ReduceOp.SUM, group=group, async_op=True)]

    @staticmethod
    def needed(tensors):
        return max(8, sum((t.numel() + 7) // 8 * 8 for t in tensors if t.dtype != torch.bfloat16))

    def wait(self):
        for w in self.works:
            w.wait()

    def grad_view(self, g):
        """The all-reduced bf16 values of the gradient view `g`: itself when it already is bf16 (reduced in place), else
        its image in the staging buffer (g lies inside one of the bucket's fp32 buffers)."""
        if g.dtype == torch.bfloat16:
            return g
        a = g.data_ptr()
        for lo, hi, o in self.spans:
            if lo <= a < hi:
                e = (a - lo) // 4
                return self.stage[o + e:o + e + g.numel()].view(g.shape)
        return None


class GradSync(object):
    def __init__(self, model, group=None, overlap=True, backbone_prefix="backbone."):
        self.group = group
        self.world = world_size(group)
        self.overlap = overlap
        self.pending = []       # [(work handle, tensors it reduces)] in launch order
        self.ready = []
        self._queued = False
        self.defer_wait = False  # True: the optimiser waits per bucket (take_buckets) instead of _final waiting for all
        # Layer buffers are merged until a bucket holds at least this many bytes before its all-reduce
        # is launched.  Stand-alone on 8 x B200 (tools/allreduce_bench.py): the 806 MB of MCAN-large
        # gradients take 1.98 ms as one call (712 GB/s bus bandwidth, NVLS) but 3.37 ms as 13 per-layer
        # calls of 62 MB -- per-call ramp-up dominates below ~100 MB.
        self.bucket_bytes = int(float(os.environ.get("MCAN_DP_BUCKET_MB", "192")) * 1e6)
        # Bucket threshold while the encoder half of the backward pass runs (MCAN_DP_ENC_BUCKET_MB).  Measured at
        # 2 x B200, MCAN-large: 40 MB (one all-reduce per encoder layer, a smaller exposed last exchange) 9.04 ms/step,
        # 192 MB 8.93 ms/step -- every extra NCCL launch pins SMs the persistent GEMMs of the chain want, so the
        # same threshold is used throughout.
        self.enc_bucket_bytes = int(float(os.environ.get("MCAN_DP_ENC_BUCKET_MB", "192")) * 1e6)
        self.acc = []
        self.acc_pairs = []     # (parameter, gradient tensor) of everything in `acc`, for the bucket-wise optimiser
        self.ready_pairs = []
        self.acc_bytes = 0
        self.early = None       # optim.EarlyStep: updates the parameters of finished buckets next to the encoder backward
        self.compress = os.environ.get("MCAN_DP_COMPRESS", "")      # "" (fp32 exchange) | "bf16"
        self._stages = []       # persistent bf16 staging buffers, one per bucket of a step (bf16 exchange)
        self._bucket_no = 0
        self.launches = 0
        self.hooks = []
        self._backwards = 0     # backward passes since the last optimiser step (overlap mode allows exactly one)
        # Only MCA_ED's backward hands its gradients over layer by layer (autograd._MCAEDRunner -> layer_hook);
        # every other backbone (MCAClassifier, stand-alone SA / SGA stacks) is reduced through parameter hooks.
        net = model.module if hasattr(model, "module") else model
        backbone = getattr(net, backbone_prefix.rstrip("."), None)
        self.layerwise = bool(getattr(backbone, "reports_layers", False))       # set by core/model/mca.py MCA_ED
        if self.world > 1 and overlap:
            for name, p in model.named_parameters():
                if p.requires_grad and not (self.layerwise and name.startswith(backbone_prefix)):
                    self.hooks.append(p.register_post_accumulate_grad_hook(self._param_ready))

    # -- overlap mode ---------------------------------------------------------------------
    def _param_ready(self, p):
        self.ready.append(p.grad)
        self.ready_pairs.append((p, p.grad))
        self._ensure_final_callback()

    def on_bufs(self, bufs, grads=None, kind=None):
        """Called inside MCA_ED.backward with the flat fp32 gradient buffers of one layer (and, for the bucket-wise
        optimiser, the {parameter: gradient view} map of that layer; kind = "dec" | "kv" | "enc")."""
        self.acc += self.ready + list(bufs)
        # (fresh view objects: holding the layer's own gradient tensors would raise their reference count and make
        #  autograd's AccumulateGrad CLONE them into .grad instead of adopting the arena views being reduced in place)
        self.acc_pairs += self.ready_pairs + ([(p, g.detach()) for p, g in grads.items() if g is not None] if grads else [])
        self.ready, self.ready_pairs = [], []
        self.acc_bytes = sum(t.numel() * 4 for t in self.acc)      # fp32-equivalent bytes: same buckets whatever the dtype
        # "enc_last": the backbone is done -- whatever has accumulated goes out now, next to the LSTM / image-projection
        # backward that follows, so that only those few gradients are left for the exposed exchange after the backward pass
        if (kind == "enc_last" and not os.environ.get("MCAN_DP_NO_LAST_FLUSH")) or self.acc_bytes >= (self.bucket_bytes if kind in (None, "dec")
                                                    else min(self.bucket_bytes, self.enc_bucket_bytes)):
            self._flush_acc()
        self._ensure_final_callback()
        if self.early is not None:
            self.early.on_dp_layer(self, kind)

    def _flush_acc(self):
        if self.acc:
            self._launch(self.acc, self.acc_pairs)
            self.acc, self.acc_pairs, self.acc_bytes = [], [], 0

    def _flush_ready(self):
        # end of backward: whatever is still waiting goes out as the last bucket
        self.acc += self.ready
        self.acc_pairs += self.ready_pairs
        self.ready, self.ready_pairs = [], []
        self._flush_acc()

    def _launch(self, tensors, pairs=None):
        if self.world == 1 or not tensors:
            return
        pairs = list(pairs) if pairs else []
        if self.compress == "bf16" and all(t.is_contiguous() for t in tensors) and \
                (self.defer_wait or all(t.dtype == torch.float32 for t in tensors)):
            if self.defer_wait and tensors[0].is_cuda:
                # the bucket-wise fused optimiser reads the reduced bf16 values in place: no unpack
                k, need = self._bucket_no, _Bf16Bucket.needed(tensors)
                while len(self._stages) <= k:
                    self._stages.append(None)
                if self._stages[k] is None or self._stages[k].numel() < need:
                    self._stages[k] = torch.empty(need, dtype=torch.bfloat16, device=tensors[0].device)
                self._bucket_no += 1
                self.pending.append((_Bf16Bucket(tensors, self.group, self._stages[k]), list(tensors), pairs))
            else:
                self.pending.append((_CompressedWork(tensors, self.group), list(tensors), pairs))
            self.launches += 1
            return
        works = _all_reduce_sum_async(tensors, self.group)
        if len(works) == len(tensors):      # one handle per tensor (backends without coalescing)
            self.pending.extend((w, [t], pairs if i == len(works) - 1 else []) for i, (w, t) in enumerate(zip(works, tensors)))
        else:
            self.pending.extend((w, tensors, pairs) for w in works)
        self.launches += 1

    def _ensure_final_callback(self):
        if not self._queued:
            Variable._execution_engine.queue_callback(self._final)
            self._queued = True

    def _final(self):
        self._flush_ready()
        self._queued = False
        self._backwards += 1
        if self._backwards > 1 and self.world > 1:
            # a second backward before the optimiser step would make autograd run `p.grad += new` on gradients that
            # are being all-reduced in place (and reduce already-reduced sums again): refuse instead of racing
            self._backwards = 0
            raise RuntimeError("mcan dp: overlap mode needs exactly one backward per optimiser step "
                               "(grad_accu_steps == 1); use dp.attach(model, overlap=False) for gradient accumulation")
        if self.defer_wait:
            return            # the optimiser consumes the buckets one by one (take_buckets)
        for entry in self.pending:
            entry[0].wait()   # current stream waits for the NCCL stream; no host sync
        self.pending = []

    def take_buckets(self):
        """[(work, tensors, [(parameter, gradient)])] of this backward in launch order; the caller waits on each work before
        it touches that bucket's gradients (FusedAdamW.step_buckets: the parameter update of the
        first buckets overlaps the all-reduce of the last ones)."""
        out, self.pending = self.pending, []
        self._backwards = 0
        self._bucket_no = 0
        return out

    def step_done(self):
        """The optimiser consumed this step's gradients (called by the overlay WarmupOptimizer / Trainer)."""
        self._backwards = 0

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


def attach(model, group=None, overlap=True, broadcast=True):
    """Makes `model` data parallel across the default (or given) process group."""
    global _active
    if _active is not None:
        _active.remove()
    _active = GradSync(model, group=group, overlap=overlap)
    if broadcast and _active.world > 1:
        broadcast_params(model, group)
    return _active


def detach():
    global _active
    if _active is not None:
        _active.remove()
    _active = None


def active():
    return _active


def layer_hook():
    """The per-layer callback MCA_ED's backward should use (None when not data parallel)."""
    if _active is not None and _active.world > 1 and _active.overlap:
        return _active.on_bufs
    return None


def broadcast_params(model, group=None, src=0):
    with torch.no_grad():
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)
        for b in model.buffers():
            dist.broadcast(b.data, src=src, group=group)


_presynced = [False]


def sync_all_grads(params, group=None, before_clip=False):
    """At-step mode: all-reduce(SUM) every existing .grad once (coalesced) and wait.

    before_clip=True: the training loop calls this itself BEFORE clip_grad_norm_ (core/exec.py:186-191 clips before
    optim.step()), so that the clip and the logged norms see the summed gradient; the overlay WarmupOptimizer.step()
    then skips its own reduction for this step."""
    if world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    for w in _all_reduce_sum_async(grads, group):
        w.wait()
    if before_clip:
        _presynced[0] = True
    return len(grads)


def consume_presync():
    """True (once) when the gradients of this step were already reduced by sync_all_grads(..., before_clip=True)."""
    done, _presynced[0] = _presynced[0], False
    return done
