"""How does the persistent GEMM (static vs dynamic tile schedule) behave when another kernel (here: a dummy that
pins N SMs, standing in for an overlapping NCCL all-reduce) occupies part of the GPU?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import capi, ops  # noqa: E402

lib = capi.load()
m, n, k = 6400, 4096, 1024
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
side = torch.cuda.Stream()


def run(hog_ctas, reps=10, limit=0):
    ops.set_sm_limit(limit)
    for _ in range(3):
        ops.gemm(a, b, out_bf16=out)
    torch.cuda.synchronize()
    if hog_ctas:
        with torch.cuda.stream(side):
            capi.check(lib.mcan_debug_hog(hog_ctas, int(2e9 * 0.004), 100 * 1024, side.cuda_stream), "hog")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        ops.gemm(a, b, out_bf16=out)
    e.record()
    torch.cuda.synchronize()
    ops.set_sm_limit(0)
    return s.elapsed_time(e) / reps * 1e3


for dynamic in (False, True):
    ops.set_gemm_schedule(dynamic)
    print("== %s tile schedule" % ("dynamic" if dynamic else "static"))
    print("GEMM 6400x4096x1024 alone:                %7.1f us" % run(0))
    for h in (4, 8, 16, 32):
        print("with %2d SMs pinned by another kernel:      %7.1f us" % (h, run(h)))
        print("   same, GEMM limited to %3d SMs:          %7.1f us" % (148 - 2 * h, run(h, limit=148 - 2 * h)))
