"""Training-step harness around the overlay Net: what core/exec.py's inner loop does
(core/exec.py:157-208: zero_grad -> forward -> BCELoss(sum) -> backward -> optimizer step),
without its per-step host syncs (loss.item(), 275 x grad-norm .cpu()), optionally captured in a
CUDA graph, and data parallel with the overlapped gradient all-reduce of dp.py.

Used by bench.py, __graft_entry__.smoke() and the loss-curve tests.  The optimiser is the
reference's: AdamW(lr=0, weight_decay=1e-4) driven by the WarmupOptimizer schedule
(core/model/optim.py:11-67).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from . import blocks, dp, ops  # noqa: E402
from . import optim as _optim  # noqa: E402
from .optim import EarlyStep, FusedAdamW  # noqa: E402


def warmup_rate(step, lr_base, data_size, batch_size):
    """WarmupOptimizer.rate (reference core/model/optim.py:36-49)."""
    per_epoch = data_size / batch_size
    for k in (1, 2, 3):
        if step <= int(per_epoch * k):
            return float(lr_base) * 0.25 * k
    return float(lr_base)


class Trainer(object):
    def __init__(self, cfg, token_size, answer_size, device, lr_base=1e-4, data_size=64 * 1000,
                 batch_size=64, state_dict=None, use_graph=False, data_parallel=False, net_cls=None,
                 fused_optimizer=True):
        from core.model.net import Net
        self.device = device
        self.cfg = cfg
        net_cls = Net if net_cls is None else net_cls
        self.net = net_cls(cfg, None, token_size, answer_size)
        if state_dict is not None:
            self.net.load_state_dict(state_dict, strict=True)
        self.net.to(device).train()
        self.lr_base, self.data_size, self.batch_size = lr_base, data_size, batch_size
        self._step = 0
        self.use_graph = use_graph
        self.lr_t = torch.zeros((), dtype=torch.float32, device=device)
        params = [p for p in self.net.parameters() if p.requires_grad]
        if fused_optimizer:
            # the library's multi-tensor AdamW: one kernel per step that also re-emits the bf16 operand copies
            self.opt = FusedAdamW(params, lr=self.lr_t if use_graph else 0.0, weight_decay=1e-4)
            self.opt.attach_shadows(self.net.all_lps())
        else:
            self.opt = torch.optim.AdamW(params, lr=self.lr_t if use_graph else 0.0, weight_decay=1e-4,
                                         fused=True, capturable=use_graph)
        self.loss_fn = torch.nn.BCELoss(reduction="sum")
        self.fused_loss = hasattr(self.net, "forward_with_loss") and os.environ.get("MCAN_FUSED_LOSS", "1") != "0"
        self.sync = dp.attach(self.net, overlap=True) if data_parallel else None
        # data parallel + fused optimiser: update bucket by bucket as the all-reduces finish
        self.bucketed = self.sync is not None and self.sync.world > 1 and isinstance(self.opt, FusedAdamW)
        if self.bucketed:
            self.sync.defer_wait = True
            # bf16 gradient exchange by default on this path (the fused optimiser reads the reduced bf16 sums in place;
            # tests/test_dp_gpu.py: 100-step loss curve within 2e-2 of the fp32 exchange, replicas bit-identical);
            # MCAN_DP_COMPRESS=fp32 keeps the fp32 exchange
            if os.environ.get("MCAN_DP_COMPRESS", "") == "":
                self.sync.compress = "bf16"
            elif os.environ["MCAN_DP_COMPRESS"] == "fp32":
                self.sync.compress = ""
            # ... and the layers' weight gradients are produced as bf16 directly in the exchange buffers
            blocks.WGRAD_BF16 = self.sync.compress == "bf16" and os.environ.get("MCAN_DP_WGRAD_BF16", "1") != "0"
        # single GPU: the optimiser update of finished layers overlaps the encoder half of the backward pass
        self.early = None
        if (isinstance(self.opt, FusedAdamW) and not self.bucketed and (self.sync is None or self.sync.world == 1)
                and os.environ.get("MCAN_EARLY_STEP", "1") != "0"):
            self.early = EarlyStep(self.opt)
            _optim.set_early(self.early)
        elif self.bucketed and os.environ.get("MCAN_EARLY_STEP", "1") != "0":
            # data parallel: finished buckets are applied next to the encoder backward as their all-reduces complete
            self.early = EarlyStep(self.opt)
            _optim.set_early(self.early)
            self.sync.early = self.early
        # kernels of the step run on a HIGH-priority stream so that they win SMs from the overlapped update
        self.main_stream = torch.cuda.Stream(device=device, priority=-1)
        self.graph = None
        self.static = None
        self.loss = None
        # device-side dropout seed word (updated inside the captured graph)
        self.ctr = torch.zeros((), dtype=torch.int64, device=device)
        self.seed_word = torch.zeros(1, dtype=torch.int32, device=device)
        if use_graph:
            ops.set_seed_tensor(self.seed_word)

    # -- one optimisation step on device-resident tensors -------------------------------------
    def _advance_seed(self):
        self.ctr.add_(1)
        x = self.ctr * 0x9E3779B97F4A7C1 + 0x7F4A7C15
        x = x ^ (x >> 29)
        x = x * 0xBF58476D1CE4E5B
        x = x ^ (x >> 32)
        self.seed_word.copy_((x & 0x7FFFFFFF).to(torch.int32).reshape(1))

    def _raw_step(self, img, ques, ans):
        self.opt.zero_grad(set_to_none=True)
        if self.use_graph:
            self._advance_seed()
        if self.fused_loss:
            loss = self.net.forward_with_loss(img, ques, ans)[0]     # sigmoid + BCE(sum) as kernels of the library
        else:
            loss = self.loss_fn(self.net(img, ques)[0], ans)
        if self.early is not None:
            self.early.begin()
        loss.backward()
        if self.bucketed:
            self.opt.step_buckets(self.sync.take_buckets())
        else:
            self.opt.step()
            if self.sync is not None:
                self.sync.step_done()
        return loss.detach()

    def _set_lr(self):
        self._step += 1
        rate = warmup_rate(self._step, self.lr_base, self.data_size, self.batch_size)
        if self.use_graph:
            self.lr_t.fill_(rate)
        else:
            for g in self.opt.param_groups:
                g["lr"] = rate

    def capture(self, img, ques, ans, warmup=3):
        """Warm up on a side stream, then capture one full training step."""
        self.static = (img.clone(), ques.clone(), ans.clone())
        s = self.main_stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._set_lr()
                self._raw_step(*self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._set_lr()
        self.opt.zero_grad(set_to_none=True)
        # thread_local: the NCCL watchdog thread polls CUDA events while we capture (data parallel)
        with torch.cuda.graph(self.graph, stream=self.main_stream, capture_error_mode="thread_local"):
            self.loss = self._raw_step(*self.static)
        self._step -= 1          # capturing records the step, it does not execute it
        torch.cuda.synchronize()

    def step(self, img, ques, ans):
        """img fp32 [B,P,I], ques int64 [B,T], ans fp32 [B,A] on the device.  Returns the loss tensor."""
        self._set_lr()
        if self.graph is not None:
            if img.data_ptr() != self.static[0].data_ptr():
                self.static[0].copy_(img, non_blocking=True)
                self.static[1].copy_(ques, non_blocking=True)
                self.static[2].copy_(ans, non_blocking=True)
            self.graph.replay()
            if isinstance(self.opt, FusedAdamW):
                self.opt.note_replay()
            return self.loss
        return self._raw_step(img, ques, ans)

    def close(self):
        if self.use_graph:
            ops.set_seed_tensor(None)
        if self.sync is not None:
            dp.detach()
            blocks.WGRAD_BF16 = False
        if self.early is not None:
            _optim.set_early(None)
