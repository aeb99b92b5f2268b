"""Runs one mcan_gemm configuration a few times (for ncu).  usage: one_gemm.py M N K cg bn [out: bf16|f32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402

m, n, k, cg, bn = (int(v) for v in sys.argv[1:6])
kind = sys.argv[6] if len(sys.argv) > 6 else "bf16"
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
b = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16 if kind == "bf16" else torch.float32)
for _ in range(3):
    if kind == "bf16":
        ops.gemm(a, b, out_bf16=out, cta_group=cg, block_n=bn)
    else:
        ops.gemm(a, b, out_f32=out, cta_group=cg, block_n=bn)
torch.cuda.synchronize()
print("done")
