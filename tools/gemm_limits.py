"""What bounds the GEMM main loop?  Times one shape with parts of the kernel switched off through
MCAN_GEMM_DEBUG (bit0: no global stores, bit1: no TMA loads, bit2: no MMAs).  Scratch tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402
from gemm_bench import timeit  # noqa: E402

shapes = [("ffn1 fwd", 6400, 4096, 1024), ("ffn2 fwd", 6400, 1024, 4096), ("merge fwd", 6400, 1024, 1024),
          ("big", 8192, 8192, 8192)]
for name, m, n, k in shapes:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for dbg, what in ((0, "normal"), (1, "no stores"), (2, "no TMA"), (3, "no TMA, no stores"), (4, "no MMA"),
                      (6, "no TMA, no MMA")):
        os.environ["MCAN_GEMM_DEBUG"] = str(dbg)
        t = timeit(lambda: ops.gemm(a, b, out_bf16=out, out_f32=None), iters=10)
        print("%-10s %-20s %8.1f us  %7.1f TFLOP/s-equivalent" % (name, what, t * 1e6, 2.0 * m * n * k / t / 1e12), flush=True)
    os.environ["MCAN_GEMM_DEBUG"] = "0"
    t = timeit(lambda: torch.matmul(a, b.t()), iters=10)
    print("%-10s %-20s %8.1f us  %7.1f TFLOP/s" % (name, "cuBLAS", t * 1e6, 2.0 * m * n * k / t / 1e12), flush=True)
