"""Data-parallel equivalence ON HARDWARE: R ranks x 64 samples with the NCCL gradient
all-reduce(SUM) of dp.py == one process on the global batch (reference semantic: nn.DataParallel + sum-reduced
BCE, core/exec.py:62-67).  Needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")]


def _worker(rank, world, port, model, out):
    import mcan_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from core.model.net import Net
    from mcan_vqa_b200 import dp
    cfg = orc.Cfg(dropout_rate=0.0, **(orc.SMALL if model == "small" else orc.LARGE))
    T, A, B = 20000, 3129, 64
    torch.manual_seed(0)
    net = Net(cfg, None, T, A).cuda().train()
    v, q, a = (t.cuda() for t in orc.synth_batch(cfg, B * world, 100, 14, T, A, seed=4321, ragged="prefix"))
    sl = slice(rank * B, (rank + 1) * B)
    loss_fn = torch.nn.BCELoss(reduction="sum")
    sync = dp.attach(net, overlap=True)
    assert sync.layerwise
    loss = loss_fn(net(v[sl], q[sl])[0], a[sl])
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.clone() for n, p in net.named_parameters()}
    launches = sync.launches
    dp.detach()
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    if rank == 0:
        # the single-process gradient of the global batch, same weights
        net.zero_grad(set_to_none=True)
        full_loss = loss_fn(net(v, q)[0], a)
        full_loss.backward()
        torch.cuda.synchronize()
        worst = (0.0, "")
        total = torch.sqrt(sum(p.grad.double().pow(2).sum() for p in net.parameters())).item()
        for n, p in net.named_parameters():
            ref = p.grad
            if ref.norm().item() < 1e-6 * total:      # mathematically zero (linear_k biases: softmax shift invariance)
                assert grads[n].norm().item() < 1e-5 * total, n
                continue
            err = ((grads[n] - ref).norm() / ref.norm()).item()
            worst = max(worst, (err, n))
        with open(out, "w") as f:
            f.write("%r\n" % ({"model": model, "world": world, "allreduce_launches": launches, "worst_rel_l2": worst,
                               "loss_sum_of_ranks": tot.item(), "loss_global": full_loss.item()},))
        assert abs(tot.item() - full_loss.item()) < 1e-3 * abs(full_loss.item())
        # two bf16 evaluations of the same gradient (64-row shards vs the 128-row batch pick different tiles / K splits)
        assert worst[0] < 4e-2, worst
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("model", ["small", "large"])
def test_nccl_ranks_times_64_equal_single_process_global_batch(model, tmp_path):
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "dp.txt")
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 1000, model, out), nprocs=world, join=True)
    text = open(out).read()
    print("DP-EQUIVALENCE", text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dp_equivalence.txt"), "a") as f:
        f.write(text)


def _curve_worker(rank, world, port, out):
    """100 optimiser steps of MCAN-small on `world` ranks x 64 samples: fp32 gradient exchange vs the bf16 exchange
    (MCAN_DP_COMPRESS=bf16: cast kernel -> bf16 all-reduce -> AdamW reads the bf16 sums in place)."""
    import mcan_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mcan_vqa_b200.train import Trainer
    cfg = orc.Cfg(dropout_rate=0.0, **orc.SMALL)
    T, A, B = 2000, 3129, 64
    dev = torch.device("cuda", rank)
    sd = orc.synth_state_dict(cfg, T, A, seed=0)
    batches = [tuple(t.cuda() for t in orc.synth_batch(cfg, B, 100, 14, T, A, seed=100 + 7 * s + rank, ragged="prefix"))
               for s in range(6)]
    curves = {}
    for mode in ("fp32", "bf16"):
        os.environ["MCAN_DP_COMPRESS"] = mode
        tr = Trainer(cfg, T, A, dev, lr_base=1e-4, data_size=25 * 64 * world, batch_size=64 * world,
                     state_dict={k: v.float() for k, v in sd.items()}, data_parallel=True)
        assert tr.bucketed
        losses = []
        for s in range(100):
            loss = tr.step(*batches[s % len(batches)]).clone()
            dist.all_reduce(loss)
            losses.append(loss.item())
        # replicas must stay identical whatever the exchange precision
        chk = torch.stack([p.detach().double().sum() for p in tr.net.parameters()])
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        assert float((hi - lo).abs().max()) == 0.0
        curves[mode] = losses
        tr.close()
    os.environ["MCAN_DP_COMPRESS"] = ""
    if rank == 0:
        worst = max(abs(a - b) / abs(a) for a, b in zip(curves["fp32"], curves["bf16"]))
        with open(out, "w") as f:
            f.write("%r\n" % ({"world": world, "steps": 100, "worst_rel_loss_diff_bf16_vs_fp32_exchange": worst,
                               "first": (curves["fp32"][0], curves["bf16"][0]), "last": (curves["fp32"][-1], curves["bf16"][-1])},))
        assert curves["fp32"][-1] < curves["fp32"][0]                  # it trains
        assert worst < 2e-2, worst
    dist.barrier()
    dist.destroy_process_group()


def test_bf16_gradient_exchange_loss_curve_matches_fp32_exchange(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "curve.txt")
    mp.spawn(_curve_worker, args=(world, 29800 + os.getpid() % 1000, out), nprocs=world, join=True)
    text = open(out).read()
    print("DP-BF16-EXCHANGE", text)
    with open(os.path.join(ROOT, "gpurun_out", "dp_equivalence.txt"), "a") as f:
        f.write(text)
