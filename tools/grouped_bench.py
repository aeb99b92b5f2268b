"""Weight-gradient GEMMs of one MCAN layer: one launch each vs one grouped launch (L2 flushed, CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=15):
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


CASES = {
    "decoder layer, large (K 6400)": (6400, [(1024, 1024), (3072, 1024), (1024, 1024), (1024, 1024), (4096, 1024), (1024, 4096)]),
    "encoder layer, large (K 896)": (896, [(1024, 4096), (4096, 1024), (3072, 1024), (1024, 1024)]),
    "decoder layer, small (K 6400)": (6400, [(512, 512), (1536, 512), (512, 512), (512, 512), (2048, 512), (512, 2048)]),
    "encoder layer, small (K 896)": (896, [(512, 2048), (2048, 512), (1536, 512), (512, 512)]),
}
for name, (rows, shapes) in CASES.items():
    probs = []
    for n, k in shapes:
        dy = (torch.randn(rows, n, device="cuda") * 0.1).to(torch.bfloat16)
        x = torch.randn(rows, k, device="cuda").to(torch.bfloat16)
        probs.append((dy, x, torch.zeros(n, k, device="cuda")))
    flops = sum(2.0 * rows * n * k for n, k in shapes)

    def single():
        for dy, x, out in probs:
            ops.gemm(dy, x, a_layout=1, b_layout=1, out_f32=out, accumulate=True)

    t1 = timeit(single)
    t2 = timeit(lambda: ops.gemm_grouped(probs))
    print("%-32s %d launches %7.1f us %7.1f TFLOP/s | grouped %7.1f us %7.1f TFLOP/s" %
          (name, len(shapes), t1 * 1e6, flops / t1 / 1e12, t2 * 1e6, flops / t2 / 1e12), flush=True)
    for s in (1, 2, 3, 4, 6):
        t = timeit(lambda: ops.gemm_grouped(probs, split_k=s))
        print("    split_k=%d %7.1f us %7.1f TFLOP/s" % (s, t * 1e6, flops / t / 1e12), flush=True)
