"""The 64-row GEMMs of the head / AttFlat merge (batch 64): auto vs forced tiles and K splits."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=12):
    ts = []
    for _ in range(iters + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


def pad8(n):
    return (n + 7) // 8 * 8


CASES = [("proj fwd", 64, 3129, 2048, 0, 0), ("proj dgrad", 64, 2048, 3129, 0, 1), ("proj wgrad", 3129, 2048, 64, 1, 1),
         ("attflat merge fwd", 64, 2048, 1024, 0, 0), ("attflat merge dgrad", 64, 1024, 2048, 0, 1),
         ("attflat merge wgrad", 2048, 1024, 64, 1, 1)]
for name, m, n, k, al, bl in CASES:
    a = (torch.randn(k, pad8(m), device="cuda")[:, :m] if al else torch.randn(m, pad8(k), device="cuda")[:, :k]).to(torch.bfloat16)
    b = (torch.randn(k, pad8(n), device="cuda")[:, :n] if bl else torch.randn(n, pad8(k), device="cuda")[:, :k]).to(torch.bfloat16) * 0.05
    out = torch.zeros(m, (n + 3) // 4 * 4, device="cuda")[:, :n]
    line = "%-20s %5dx%5dx%5d |" % (name, m, n, k)
    best = (1e9, "")
    for acc in (False, True):
        for bn in (0, 64, 128, 256):
            for sk in ((0,) if not acc else (0, 2, 4, 8, 16)):
                if acc and bn == 64:
                    continue
                try:
                    t = timeit(lambda: ops.gemm(a, b, a_layout=al, b_layout=bl, out_f32=out, accumulate=acc, split_k=sk,
                                                block_n=bn, cta_group=1 if bn else 0))
                except Exception:  # noqa: BLE001
                    continue
                tag = "%s/bn%d%s" % ("acc" if acc else "st", bn, ("/s%d" % sk) if sk else "")
                line += " %s %.1f" % (tag, t * 1e6)
                if t < best[0]:
                    best = (t, tag)
    print(line + " || best %s %.1f us" % (best[1], best[0] * 1e6), flush=True)
