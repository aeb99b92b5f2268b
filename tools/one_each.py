"""Launches every memory-bound kernel of the hot path a few times on its MCAN-large image-side shape
(batch 64: 6400 rows x 1024, 16 heads of 64, 100 regions; AdamW on 67 M parameters) -- the program ncu
profiles for the achieved-HBM-bandwidth evidence in profiles/ (one --set full capture per kernel):

    ncu --set full -k regex:'ln_|attn_|adamw' -s 10 -c 5 python tools/one_each.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402
from mcan_vqa_b200.optim import FusedAdamW  # noqa: E402

rows, H, B, heads, d, S = 6400, 1024, 64, 16, 64, 100
dev = "cuda"
x = torch.randn(rows, H, device=dev)
a2, b2 = torch.rand(H, device=dev) + 0.5, torch.randn(H, device=dev)
y32 = torch.empty_like(x)
ybf = torch.empty(rows, H, device=dev, dtype=torch.bfloat16)
mean, sigma = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
dy = torch.randn(rows, H, device=dev)
dx = torch.empty_like(x)
dxbf = torch.empty_like(ybf)
da, db, dbias = (torch.zeros(H, device=dev) for _ in range(3))
qkv = torch.randn(rows, 3 * H, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
mask = torch.zeros(B, S, dtype=torch.uint8, device=dev)
out = torch.empty(rows, H, device=dev, dtype=torch.bfloat16)
do = torch.randn(rows, H, device=dev).to(torch.bfloat16)
dqkv = torch.empty(rows, 3 * H, device=dev, dtype=torch.bfloat16)
kw = dict(batch=B, heads=heads, sq=S, sk=S, head_dim=d, scale=1.0 / math.sqrt(d), dropout_p=0.1, seed=7)
params = [torch.nn.Parameter(torch.randn(4096, 1024, device=dev)) for _ in range(16)]      # 67 M parameters
shadow = [torch.empty(4096, 1024, device=dev, dtype=torch.bfloat16) for _ in params]
opt = FusedAdamW(params, lr=1e-4, weight_decay=1e-4)
for p, s in zip(params, shadow):
    opt._shadow[id(p)] = [s]
    p.grad = torch.randn_like(p)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

for _ in range(3):      # 5 library kernels per round, in this order
    flush.zero_()
    ops.layernorm_fwd(x, a2, b2, 1e-6, y_f32=y32, y_bf16=ybf, mean=mean, sigma=sigma)
    flush.zero_()
    ops.layernorm_bwd(dy, x, mean, sigma, a2, 1e-6, dx_f32=dx, dx_bf16=dxbf, dropout_p=0.1, seed=3, da2=da, db2=db,
                      dbias=dbias)
    flush.zero_()
    ops.attn_fwd(q, k, v, mask, out, **kw)
    flush.zero_()
    ops.attn_bwd(q, k, v, mask, do, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], **kw)
    flush.zero_()
    opt.step()
torch.cuda.synchronize()
print("done")
