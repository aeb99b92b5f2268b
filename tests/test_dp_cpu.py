"""Data-parallel gradient exchange (mcan-vqa_b200/dp.py) on 2 gloo ranks, CPU only.

Checks the reference's DataParallel semantics we replace (core/exec.py:62-67): per-rank
sum-reduced losses + all-reduce(SUM) == single-process gradient of the global batch, for both the
overlapped mode (layer buffers + post-accumulate hooks + end-of-backward wait) and the at-step
mode used by the overlay WarmupOptimizer."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _LayerFn(torch.autograd.Function):
    """Stand-in for MCA_ED's backward: produces parameter grads in flat buffers and hands them to dp."""

    @staticmethod
    def forward(ctx, x, w1, w2):
        ctx.save_for_backward(x, w1, w2)
        return (x * w1).sum(1, keepdim=True) * w2

    @staticmethod
    def backward(ctx, g):
        from mcan_vqa_b200 import dp
        x, w1, w2 = ctx.saved_tensors
        s = (x * w1).sum(1, keepdim=True)
        flat = torch.zeros(w1.numel() + w2.numel())
        g1 = flat[: w1.numel()].view_as(w1)
        g2 = flat[w1.numel():].view_as(w2)
        g1.copy_(((g * w2) * x).sum(0))
        g2.copy_((g * s).sum(0))
        hook = dp.layer_hook() if _Toy.use_hook else None
        if hook is not None:
            hook([flat])
        return (g * w2) * w1, g1, g2


class _Toy(torch.nn.Module):
    use_hook = True

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.pre = torch.nn.Linear(6, 5)
        self.backbone = torch.nn.Module()
        self.backbone.reports_layers = True      # like MCA_ED: its backward calls dp.layer_hook()
        self.backbone.w1 = torch.nn.Parameter(torch.randn(5, generator=g))
        self.backbone.w2 = torch.nn.Parameter(torch.randn(1, generator=g))
        self.post = torch.nn.Linear(1, 3)

    def forward(self, x):
        h = _LayerFn.apply(self.pre(x), self.backbone.w1, self.backbone.w2)
        return torch.sigmoid(self.post(h))


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcan_vqa_b200 import dp
    torch.manual_seed(0)
    model = _Toy()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 6, generator=g)
    y = torch.rand(8, 3, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    loss_fn = torch.nn.BCELoss(reduction="sum")
    if mode == "overlap":
        sync = dp.attach(model, overlap=True)
        sync.bucket_bytes = 0          # every layer's buffers go out as soon as they are complete
        loss_fn(model(xs), ys).backward()
        assert sync.launches >= 2      # layer buffers + hooked parameters went out separately
    elif mode == "merged":
        sync = dp.attach(model, overlap=True)
        assert sync.bucket_bytes > 1e6  # default: layers are merged into >= 192 MB buckets
        loss_fn(model(xs), ys).backward()
        assert sync.launches == 1      # the toy model's gradients fit one bucket
    elif mode == "bf16":
        # opt-in compressed exchange: bf16 on the wire, fp32 sums written back into .grad
        sync = dp.attach(model, overlap=True)
        sync.compress = "bf16"
        loss_fn(model(xs), ys).backward()
        assert sync.launches == 1
    elif mode == "buckets":
        # the optimiser consumes the all-reduces one by one (FusedAdamW.step_buckets protocol):
        # nothing is waited for at the end of backward, every gradient belongs to exactly one bucket
        sync = dp.attach(model, overlap=True)
        sync.defer_wait = True
        sync.bucket_bytes = 0
        loss_fn(model(xs), ys).backward()
        buckets = sync.take_buckets()
        assert len(buckets) >= 2 and sync.pending == []
        seen = set()
        for work, tensors, _pairs in buckets:
            work.wait()
            for t in tensors:
                lo, hi = t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()
                for n, p in model.named_parameters():
                    if lo <= p.grad.data_ptr() < hi:
                        assert n not in seen
                        seen.add(n)
        assert seen == {n for n, _ in model.named_parameters()}
    elif mode == "plain_backbone":
        # a backbone whose backward does NOT report its layers (MCAClassifier, stand-alone SA / SGA stacks):
        # every parameter, backbone included, must be reduced through the parameter hooks
        model.backbone.reports_layers = False
        sync = dp.attach(model, overlap=True)
        assert not sync.layerwise and len(sync.hooks) == len(list(model.parameters()))
        _Toy.use_hook = False
        loss_fn(model(xs), ys).backward()
    elif mode == "second_backward":
        # overlap mode reduces gradient buffers in place: a second backward before the optimiser step must be refused
        sync = dp.attach(model, overlap=True)
        loss_fn(model(xs), ys).backward()
        with pytest.raises(RuntimeError, match="one backward per optimiser step"):
            loss_fn(model(xs), ys).backward()
        model.zero_grad(set_to_none=True)
        sync.pending = []
        loss_fn(model(xs), ys).backward()      # after the refusal the counter restarts: a normal step works again
    elif mode == "presync":
        # the training loop reduces BEFORE clip_grad_norm_ (exec.py clips before optim.step()): step() must not reduce again
        dp.attach(model, overlap=False)
        sys.path.insert(0, ROOT)
        from core.model.optim import WarmupOptimizer
        opt = WarmupOptimizer(0.0, torch.optim.SGD(model.parameters(), lr=0.0), 64, 8)
        loss_fn(model(xs), ys).backward()
        assert dp.sync_all_grads(model.parameters(), before_clip=True) > 0
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)       # sees the summed gradient
        opt.step()
        assert not dp.consume_presync()                              # consumed by step()
    else:
        dp.attach(model, overlap=False)
        sys.path.insert(0, ROOT)
        from core.model.optim import WarmupOptimizer
        opt = WarmupOptimizer(0.0, torch.optim.SGD(model.parameters(), lr=0.0), 64, 8)
        loss_fn(model(xs), ys).backward()
        opt.step()                     # all-reduces inside step()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    dp.detach()                        # the single-process reference below must not all-reduce
    if rank == 0:
        ref = _Toy()
        ref.load_state_dict(model.state_dict())
        loss_fn(ref(x), y).backward()
        tol = dict(rtol=2e-2, atol=1e-3) if mode == "bf16" else dict(rtol=1e-5, atol=1e-6)
        for n, p in ref.named_parameters():
            assert torch.allclose(grads[n], p.grad, **tol), n
            if mode == "bf16":     # the values really went through bf16
                assert torch.equal(grads[n], grads[n].to(torch.bfloat16).float()), n
        open(out, "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["overlap", "merged", "bf16", "buckets", "at_step", "plain_backbone", "second_backward", "presync"])
def test_two_rank_sum_allreduce_equals_global_batch_gradient(mode, tmp_path):
    out = str(tmp_path / "ok.txt")
    port = 29500 + (os.getpid() % 2000) + {"overlap": 0, "at_step": 1, "buckets": 2, "merged": 3, "bf16": 4, "plain_backbone": 5, "second_backward": 6, "presync": 7}[mode]
    mp.spawn(_worker, args=(2, port, mode, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
