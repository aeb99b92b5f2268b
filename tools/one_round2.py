"""Launches the round-2 kernels once each on their MCAN-large shapes (batch 64) -- the program ncu profiles for the
evidence in profiles/r02_ncu_full_round2_kernels.metrics.csv:

    ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|lstm_|attflat_|rowmask|sigmoid' \
        python tools/one_round2.py
Order: grouped wgrad launch of a decoder layer (6 problems, K 6400), of an encoder layer (4 problems, K 896), LSTM
forward, LSTM backward, AttFlat pool forward / backward (image side), rowmask_cast, sigmoid+BCE forward / backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import blocks, ops  # noqa: E402
from mcan_vqa_b200.blocks import LinearParams, Runtime  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def group(rows, shapes):
    probs = []
    for n, k in shapes:
        dy = (torch.randn(rows, n, device=dev) * 0.1).to(torch.bfloat16)
        x = torch.randn(rows, k, device=dev).to(torch.bfloat16)
        probs.append((dy, x, torch.empty(n, k, device=dev)))
    return probs


dec = group(6400, [(1024, 1024), (3072, 1024), (1024, 1024), (1024, 1024), (4096, 1024), (1024, 4096)])
enc = group(896, [(1024, 4096), (4096, 1024), (3072, 1024), (1024, 1024)])
B, T, E, H, V = 64, 14, 300, 1024, 20000
emb = torch.nn.Embedding(V, E).cuda()
lstm = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).cuda()
tokens = torch.randint(1, V, (B, T), device=dev)
lp_ih = LinearParams([(lstm.weight_ih_l0, lstm.bias_ih_l0)], pad=64).get(True, True)
lp_hh = LinearParams([(lstm.weight_hh_l0, lstm.bias_hh_l0)]).get(True)
dq = torch.randn(B * T, H, device=dev) * 0.1
S, M, G = 100, 512, 1
hmid = torch.relu(torch.randn(B * S, M, device=dev)).to(torch.bfloat16)
w2, b2 = torch.randn(G, M, device=dev) * 0.1, torch.randn(G, device=dev)
x = torch.randn(B * S, H, device=dev)
mask = torch.zeros(B, S, dtype=torch.uint8, device=dev)
att_w = torch.empty(B, S, G, device=dev)
p32 = torch.empty(B, G * H, device=dev)
pbf = torch.empty(B, G * H, device=dev, dtype=torch.bfloat16)
dpooled = torch.randn(B, G * H, device=dev)
dx = torch.empty(B * S, H, device=dev)
dh = torch.empty(B * S, M, device=dev, dtype=torch.bfloat16)
dw2, db2 = torch.zeros(G, M, device=dev), torch.zeros(G, device=dev)
feat = torch.randn(6400, 2048, device=dev)
fbf = torch.empty(6400, 2048, device=dev, dtype=torch.bfloat16)
fmask = torch.empty(6400, dtype=torch.uint8, device=dev)
logits = torch.randn(64, 3132, device=dev)[:, :3129]
target = torch.rand(64, 3129, device=dev)
probs = torch.empty(64, 3129, device=dev)
loss = torch.zeros((), device=dev)
dz = torch.empty(64, 3136, device=dev, dtype=torch.bfloat16)[:, :3129]
dbias = torch.zeros(3129, device=dev)
kw = dict(batch=B, s=S, h=H, mlp=M, glimpses=G)
for _ in range(2):
    flush.zero_()
    ops.gemm_grouped(dec, accumulate=False)
    flush.zero_()
    ops.gemm_grouped(enc, accumulate=False)
    flush.zero_()
    rt = Runtime(True, 0.0)
    q, _, ctx = blocks.qenc_fwd(rt, emb.weight.detach(), lp_ih, lp_hh, tokens, True)
    flush.zero_()
    blocks.qenc_bwd(rt, ctx, dq, V)
    flush.zero_()
    ops.attflat_pool_fwd(hmid, w2, b2, mask, x, att_w=att_w, pooled_f32=p32, pooled_bf16=pbf, **kw)
    flush.zero_()
    ops.attflat_pool_bwd(dpooled, p32, hmid, w2, mask, x, att_w, gate_scale=1.0, dx=dx, dhmid=dh, dw2=dw2, db2=db2, **kw)
    flush.zero_()
    ops.rowmask_cast(feat, fbf, None, fmask)
    ops.sigmoid_bce_fwd(logits, probs, target, loss)
    ops.sigmoid_bce_bwd(probs, dz, target=target, dbias=dbias)
torch.cuda.synchronize()
print("done")
