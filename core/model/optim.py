"""Overlay of the reference's core/model/optim.py (WarmupOptimizer / get_optim / adjust_lr).

Same interface and schedule as /root/reference/core/model/optim.py:11-75.  The one addition:
when torch.distributed is initialised, step() first all-reduces (SUM) every gradient across the
ranks -- the one-process-per-GPU replacement of the reference's nn.DataParallel gradient
reduction (core/exec.py:62-63) that works with core/exec.py unchanged, for any grad_accu_steps.
"""
import os

import torch
from torch.optim import AdamW

from mcan_vqa_b200 import dp


class WarmupOptimizer(object):
    def __init__(self, lr_base, optimizer, data_size, batch_size):
        self.optimizer = optimizer
        self._step = 0
        self.lr_base = lr_base
        self._rate = 0
        self.data_size = data_size
        self.batch_size = batch_size

    def step(self):
        self._step += 1
        sync = dp.active()
        if sync is None or not sync.overlap:
            if not dp.consume_presync():      # (the loop may have reduced before clip_grad_norm_: dp.sync_all_grads(..., before_clip=True))
                dp.sync_all_grads([p for g in self.optimizer.param_groups for p in g['params']])
        else:
            sync.step_done()
        self._rate = self.rate()
        for group in self.optimizer.param_groups:
            group['lr'] = self._rate
        self.optimizer.step()

    def zero_grad(self):
        self.optimizer.zero_grad()

    def rate(self, step=None):
        """lr_base x {1/4, 2/4, 3/4, 1} over the first three epochs (reference optim.py:36-49)."""
        step = self._step if step is None else step
        steps_per_epoch = self.data_size / self.batch_size
        for k in (1, 2, 3):
            if step <= int(steps_per_epoch * k):
                return float(self.lr_base) * 0.25 * k
        return float(self.lr_base)


def get_optim(opt, model, data_size, lr_base=None):
    """reference optim.py:51-67.  On the GPU the AdamW is the library's fused multi-tensor kernel
    (same update rule; it also keeps the bf16 GEMM-operand copies of the weights current);
    MCAN_FUSED_ADAMW=0 or CPU parameters select torch.optim.AdamW."""
    lr_base = opt.lr_base if lr_base is None else lr_base
    params = [p for p in model.parameters() if p.requires_grad]
    inner = None
    if params and all(p.is_cuda for p in params) and os.environ.get("MCAN_FUSED_ADAMW", "1") != "0":
        from mcan_vqa_b200.optim import FusedAdamW
        inner = FusedAdamW(params, lr=0, weight_decay=1e-4)
        net = model.module if hasattr(model, "module") else model
        if hasattr(net, "all_lps"):
            inner.attach_shadows(net.all_lps())
    if inner is None:
        inner = AdamW(params, lr=0, weight_decay=1e-4)
    return WarmupOptimizer(lr_base, inner, data_size, opt.batch_size)


def adjust_lr(optim, decay_r):
    optim.lr_base = decay_r * float(optim.lr_base)


def adjust_reg_factor(factor, decay_r):
    factor *= decay_r
