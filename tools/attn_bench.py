"""Times the fused attention kernels (forward / backward) on the MCAN shapes with CUDA events and
prints achieved HBM GB/s against the algorithmic bytes (Q,K,V read + O written; backward: Q,K,V,dO
read + dQ,dK,dV written).  Scratch tool."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


for name, B, heads, d, sq, sk in (("image self  (large)", 64, 16, 64, 100, 100), ("guided      (large)", 64, 16, 64, 100, 14),
                                  ("question    (large)", 64, 16, 64, 14, 14), ("image self  (small)", 64, 8, 64, 100, 100),
                                  ("image self  d=128  ", 64, 8, 128, 100, 100)):
    H = heads * d
    qkv = torch.randn(B * sq, 3 * H, device="cuda").to(torch.bfloat16)
    kv = qkv if sk == sq else torch.randn(B * sk, 3 * H, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :H], kv[:, H:2 * H], kv[:, 2 * H:]
    mask = torch.zeros(B, sk, dtype=torch.uint8, device="cuda")
    out = torch.empty(B * sq, H, device="cuda", dtype=torch.bfloat16)
    do = torch.randn(B * sq, H, device="cuda").to(torch.bfloat16)
    dq, dk, dv = (torch.empty(B * s, H, device="cuda", dtype=torch.bfloat16) for s in (sq, sk, sk))
    kw = dict(batch=B, heads=heads, sq=sq, sk=sk, head_dim=d, scale=1.0 / math.sqrt(d), dropout_p=0.1, seed=7)
    tf = timeit(lambda: ops.attn_fwd(q, k, v, mask, out, **kw))
    tb = timeit(lambda: ops.attn_bwd(q, k, v, mask, do, dq, dk, dv, **kw))
    bf = 2.0 * H * B * (2 * sq + 2 * sk)
    bb = 2.0 * H * B * (3 * sq + 4 * sk)
    print("%s fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s" % (name, tf * 1e6, bf / tf / 1e9, tb * 1e6, bb / tb / 1e9), flush=True)
