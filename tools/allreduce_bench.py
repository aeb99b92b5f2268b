"""Stand-alone NCCL all-reduce(SUM) bandwidth on the gradient volume of one MCAN-large step
(806 MB fp32), to compare with what the exchange costs inside the training step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/allreduce_bench.py
"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
TOTAL = 201_551_291      # parameters of MCAN-large (token_size 20000)


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for name, dtype, nbuckets in (("fp32, 1 call", torch.float32, 1), ("fp32, 13 buckets", torch.float32, 13),
                              ("bf16, 1 call", torch.bfloat16, 1), ("bf16, 13 buckets", torch.bfloat16, 13)):
    n = TOTAL // nbuckets
    bufs = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(nbuckets)]

    def run():
        works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in bufs]
        for w in works:
            w.wait()

    ms = timed(run)
    nbytes = n * nbuckets * bufs[0].element_size()
    if rank == 0:
        alg = nbytes / ms / 1e6
        print("%-18s %7.1f MB  %7.3f ms  algbw %6.1f GB/s  busbw %6.1f GB/s" %
              (name, nbytes / 1e6, ms, alg, alg * 2 * (world - 1) / world), flush=True)
    del bufs
torch.cuda.synchronize()
os._exit(0)
