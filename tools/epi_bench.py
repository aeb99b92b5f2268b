"""Times the tcgen05 GEMM with each fused-epilogue stage switched on (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402
from gemm_bench import timeit  # noqa: E402


def run(name, m, n, k, **kw):
    t = timeit(lambda: ops.gemm(a, b, **kw))
    print("%-44s %8.1f us  %7.1f TFLOP/s" % (name, t * 1e6, 2.0 * m * n * k / t / 1e12), flush=True)


m, n, k = 6400, 4096, 1024
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
obf = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
o32 = torch.empty(m, n, device="cuda")
resid = torch.randn(m, n, device="cuda")
gate = torch.randn(m, n, device="cuda").to(torch.bfloat16)
run("ffn1 6400x4096x1024 bf16 out", m, n, k, out_bf16=obf)
run("  + bias", m, n, k, out_bf16=obf, bias=bias)
run("  + bias + relu", m, n, k, out_bf16=obf, bias=bias, relu=True)
run("  + bias + relu + dropout", m, n, k, out_bf16=obf, bias=bias, relu=True, dropout_p=0.1, seed=1)
run("  + gate (dgrad through relu)", m, n, k, out_bf16=obf, gate=gate, gate_scale=1.1)
run("  fp32 out", m, n, k, out_f32=o32)
run("  fp32 out + resid", m, n, k, out_f32=o32, resid=resid)
run("  fp32 out + resid + bias + dropout", m, n, k, out_f32=o32, resid=resid, bias=bias, dropout_p=0.1, seed=1)
m, n, k = 6400, 1024, 1024
b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
obf = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
o32 = torch.empty(m, n, device="cuda")
resid = torch.randn(m, n, device="cuda")
run("merge 6400x1024x1024 bf16 out", m, n, k, out_bf16=obf)
run("  fp32 out + resid + bias + dropout", m, n, k, out_f32=o32, resid=resid, bias=bias, dropout_p=0.1, seed=1)
