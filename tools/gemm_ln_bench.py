"""Fused GEMM + residual + LayerNorm (mcan_gemm_ln) vs the unfused chain (mcan_gemm -> mcan_layernorm_fwd),
CUDA events, L2 flushed between iterations.  Scratch tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402
from gemm_bench import timeit  # noqa: E402

for m, n, k in [(6400, 1024, 1024), (6400, 1024, 4096), (896, 1024, 1024), (896, 1024, 4096), (6400, 512, 512), (6400, 512, 2048)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda")
    a2, b2 = torch.ones(n, device="cuda"), torch.zeros(n, device="cuda")
    s, y32 = torch.empty(m, n, device="cuda"), torch.empty(m, n, device="cuda")
    ybf = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    mean, sigma = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")

    def fused():
        ops.gemm_ln(a, w, bias=bias, resid=resid, ln_a2=a2, ln_b2=b2, eps=1e-6, dropout_p=0.1, seed=1, s_f32=s, y_f32=y32,
                    y_bf16=ybf, mean=mean, sigma=sigma)

    def unfused():
        ops.gemm(a, w, bias=bias, dropout_p=0.1, seed=1, resid=resid, out_f32=s)
        ops.layernorm_fwd(s, a2, b2, 1e-6, y_f32=y32, y_bf16=ybf, mean=mean, sigma=sigma)

    def gemm_only():
        ops.gemm(a, w, bias=bias, dropout_p=0.1, seed=1, resid=resid, out_f32=s)

    tf, tu, tg = timeit(fused), timeit(unfused), timeit(gemm_only)
    print("%5d x %4d x %4d   fused %6.1f us   unfused %6.1f us (GEMM alone %6.1f)   fused GEMM-equivalent %6.0f TFLOP/s" %
          (m, n, k, tf * 1e6, tu * 1e6, tg * 1e6, 2.0 * m * n * k / tf / 1e12), flush=True)
