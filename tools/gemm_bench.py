"""Times mcan_gemm on the MCAN GEMM shapes with CUDA events (L2 flushed between iterations)
and prints TFLOP/s next to torch.matmul (cuBLAS) on the same shapes.  Scratch tool."""
import json
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


def main():
    res = []
    shapes = [
        # name, M, N, K, a_layout, b_layout, accumulate
        ("fwd qkv large", 6400, 3072, 1024, 0, 0, False),
        ("fwd merge large", 6400, 1024, 1024, 0, 0, False),
        ("fwd ffn1 large", 6400, 4096, 1024, 0, 0, False),
        ("fwd ffn2 large", 6400, 1024, 4096, 0, 0, False),
        ("fwd qkv small", 6400, 1536, 512, 0, 0, False),
        ("fwd ffn1 small", 6400, 2048, 512, 0, 0, False),
        ("fwd enc qkv large", 896, 3072, 1024, 0, 0, False),
        ("dgrad ffn2 large", 6400, 4096, 1024, 0, 1, False),
        ("dgrad ffn1 large", 6400, 1024, 4096, 0, 1, False),
        ("wgrad ffn1 large", 4096, 1024, 6400, 1, 1, True),
        ("wgrad merge large", 1024, 1024, 6400, 1, 1, True),
        ("wgrad merge small", 512, 512, 6400, 1, 1, True),
        ("wgrad qkv large", 3072, 1024, 6400, 1, 1, True),
        ("dgrad merge large", 6400, 1024, 1024, 0, 1, False),
        ("enc ffn2 large", 896, 1024, 4096, 0, 0, False),
    ]
    for name, m, n, k, al, bl, acc in shapes:
        a = torch.randn((m, k) if al == 0 else (k, m), device="cuda").to(torch.bfloat16)
        b = (torch.randn((n, k) if bl == 0 else (k, n), device="cuda") * 0.05).to(torch.bfloat16)
        for cg, bn in ((2, 256), (4, 256), (0, 0)):
            if acc:
                out = torch.zeros(m, n, device="cuda")
                fn = lambda: ops.gemm(a, b, a_layout=al, b_layout=bl, out_f32=out, accumulate=True, block_n=bn, cta_group=cg)
            else:
                out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
                fn = lambda: ops.gemm(a, b, a_layout=al, b_layout=bl, out_bf16=out, block_n=bn, cta_group=cg)
            t = timeit(fn)
            tf = 2.0 * m * n * k / t / 1e12
            res.append({"name": name, "cta_group": cg, "block_n": bn, "us": t * 1e6, "tflops": tf})
            print("%-20s CG=%d BN=%3d  %8.1f us  %7.1f TFLOP/s" % (name, cg, bn, t * 1e6, tf), flush=True)
        A = a if al == 0 else a.t()
        B = b.t() if bl == 0 else b
        t = timeit(lambda: torch.matmul(A, B))
        print("%-20s cuBLAS  %8.1f us  %7.1f TFLOP/s" % (name, t * 1e6, 2.0 * m * n * k / t / 1e12), flush=True)
        res.append({"name": name, "block_n": "cublas", "us": t * 1e6, "tflops": 2.0 * m * n * k / t / 1e12})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
