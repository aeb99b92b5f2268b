"""Diagnostic comparison of the attention kernels (whichever path mcan_attn_fwd / _bwd dispatch to; MCAN_ATTN_TC=0
forces the mma.sync kernels) against fp32 torch math on the MCAN shapes: prints max errors per output instead of
asserting, so that one GPU call tells which product of a new kernel is wrong.  Scratch tool."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcan_vqa_b200 import ops  # noqa: E402


def heads_view(t, batch, s, heads, d):
    return t.float().reshape(batch, s, heads, d).transpose(1, 2)


def ref_attn(qh, kh, vh, mask, scale, keep=None, p=0.0):
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :], -1e9)
    a = torch.softmax(s, dim=-1)
    if keep is not None:
        a = a * keep / (1.0 - p)
    return torch.matmul(a, vh)


def run(batch, heads, sq, sk, kind, p, seed=11):
    d = 64
    g = torch.Generator().manual_seed(sq * 131 + sk)
    H = heads * d
    qkv = torch.randn(batch * max(sq, sk), 3 * H, generator=g).to(torch.bfloat16).cuda()
    q, k, v = qkv[: batch * sq, :H], qkv[: batch * sk, H:2 * H], qkv[: batch * sk, 2 * H:]
    mask = None
    if kind == "random":
        mask = (torch.rand(batch, sk, generator=g) < 0.3)
        mask[0, :] = True
        mask = mask.cuda()
    elif kind == "prefix":
        lens = torch.randint(1, sk + 1, (batch,), generator=g)
        mask = (torch.arange(sk)[None, :] >= lens[:, None]).cuda()
    mu8 = None if mask is None else mask.to(torch.uint8).contiguous()
    scale = 1.0 / math.sqrt(d)
    kw = dict(batch=batch, heads=heads, sq=sq, sk=sk, head_dim=d, scale=scale, dropout_p=p, seed=seed)
    out = torch.full((batch * sq, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attn_fwd(q, k, v, mu8, out, **kw)
    torch.cuda.synchronize()
    keep = None
    if p > 0:
        keep = ops.dropout_keep_mask(batch * heads * sq * sk, p, seed).view(batch, heads, sq, sk).cuda().float()
    qh = heads_view(q, batch, sq, heads, d).requires_grad_(True)
    kh = heads_view(k, batch, sk, heads, d).requires_grad_(True)
    vh = heads_view(v, batch, sk, heads, d).requires_grad_(True)
    ref = ref_attn(qh, kh, vh, mask, scale, keep, p)
    got = heads_view(out, batch, sq, heads, d)
    msg = "sq %3d sk %3d %-6s p %.1f | out err %.3e (ref max %.2f, nan %d)" % (
        sq, sk, kind, p, (got - ref).abs().nan_to_num(99.0).max().item(), ref.abs().max().item(), int(torch.isnan(got).sum()))
    dout = (torch.randn(batch * sq, H, device="cuda") * 0.5).to(torch.bfloat16)
    dqkv = torch.full((batch * max(sq, sk), 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = dqkv[: batch * sq, :H], dqkv[: batch * sk, H:2 * H], dqkv[: batch * sk, 2 * H:]
    ops.attn_bwd(q, k, v, mu8, dout, dq, dk, dv, **kw)
    torch.cuda.synchronize()
    ref.backward(heads_view(dout, batch, sq, heads, d))
    for name, gt, rf, s in (("dq", dq, qh.grad, sq), ("dk", dk, kh.grad, sk), ("dv", dv, vh.grad, sk)):
        gt = heads_view(gt, batch, s, heads, d)
        msg += " | %s err %.3e (max %.2f)" % (name, (gt - rf).abs().nan_to_num(99.0).max().item(), rf.abs().max().item())
    print(msg, flush=True)


if __name__ == "__main__":
    print("MCAN_ATTN_TC =", os.environ.get("MCAN_ATTN_TC", "(default)"))
    for case in ((3, 8, 100, 100, "none", 0.0), (3, 8, 100, 100, "random", 0.0), (3, 8, 100, 100, "random", 0.1),
                 (3, 8, 100, 14, "prefix", 0.0), (3, 8, 100, 14, "prefix", 0.1), (2, 16, 60, 60, "none", 0.0),
                 (2, 4, 128, 128, "random", 0.1), (2, 4, 64, 77, "random", 0.1), (2, 4, 50, 33, "prefix", 0.1),
                 (2, 4, 99, 51, "random", 0.1)):
        try:
            run(*case)
        except Exception as ex:  # noqa: BLE001
            print("case", case, "FAILED:", repr(ex)[:300], flush=True)
            break
