"""Launches the image self-attention forward and backward once each (MCAN-large: batch 64, 16 heads, 100 x 100, head dim
64, dropout 0.1) after an L2 flush -- the program ncu profiles for profiles/r02_ncu_full_attention_tc.metrics.csv:

    ncu --set full --clock-control none --import-source on -k regex:attn_ -o gpurun_out/attn_tc python tools/one_attn.py
MCAN_ATTN_TC=0 profiles the mma.sync kernels instead."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402

B, heads, d, sq, sk = 64, 16, 64, 100, 100
H = heads * d
qkv = torch.randn(B * sq, 3 * H, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
mask = torch.zeros(B, sk, dtype=torch.uint8, device="cuda")
out = torch.empty(B * sq, H, device="cuda", dtype=torch.bfloat16)
do = torch.randn(B * sq, H, device="cuda").to(torch.bfloat16)
dqkv = torch.empty(B * sq, 3 * H, device="cuda", dtype=torch.bfloat16)
dq, dk, dv = dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:]
kw = dict(batch=B, heads=heads, sq=sq, sk=sk, head_dim=d, scale=1.0 / math.sqrt(d), dropout_p=0.1, seed=7)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    ops.attn_fwd(q, k, v, mask, out, **kw)
    flush.zero_()
    ops.attn_bwd(q, k, v, mask, do, dq, dk, dv, **kw)
torch.cuda.synchronize()
