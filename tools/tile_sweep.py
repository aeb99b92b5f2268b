"""Every distinct GEMM (shape x operand layout x epilogue) of an MCAN-large training step: the auto-picked tile /
split vs every forced alternative (L2 flushed, CUDA events) -- input for the picker rules in gemm_tcgen05.cu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=12):
    ts = []
    for _ in range(iters + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


CASES = [   # (name, m, n, k, b_layout, epilogue)
    ("dgrad ffn2", 6400, 4096, 1024, 1, "gate"), ("ffn1 fwd", 6400, 4096, 1024, 0, "relu"),
    ("ffn2 fwd", 6400, 1024, 4096, 0, "resid_drop"), ("dgrad ffn1", 6400, 1024, 4096, 1, "resid"),
    ("dgrad qkv", 6400, 1024, 3072, 1, "resid"), ("dgrad merge", 6400, 1024, 1024, 1, "bf16"),
    ("qkv fwd", 6400, 3072, 1024, 0, "bias_bf16"), ("q fwd", 6400, 1024, 1024, 0, "bias_bf16"),
    ("img linear", 6400, 1024, 2048, 0, "bias_f32"),
    ("enc dgrad ffn2", 896, 4096, 1024, 1, "gate"), ("enc ffn1", 896, 4096, 1024, 0, "relu"),
    ("enc qkv", 896, 3072, 1024, 0, "bias_bf16"), ("enc merge", 896, 1024, 1024, 0, "resid_drop"),
    ("enc dgrad merge", 896, 1024, 1024, 1, "bf16"), ("kv all", 896, 12288, 1024, 0, "bias_bf16"),
    ("enc ffn2 fwd", 896, 1024, 4096, 0, "resid_drop"),
    ("enc dgrad ffn1 (split-K)", 896, 1024, 4096, 1, "acc_resid"), ("enc dgrad qkv (split-K)", 896, 1024, 3072, 1, "acc_resid"),
    ("kv dgrad (split-K)", 896, 1024, 12288, 1, "acc_resid"),
]
for name, m, n, k, bl, epi in CASES:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    if bl:
        w = w.t().contiguous()          # [K, N]
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda")
    gate = torch.randn(m, n, device="cuda").to(torch.bfloat16)
    out = torch.zeros(m, n, device="cuda")
    outb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(n, device="cuda")
    kw = {"gate": dict(gate=gate, gate_scale=1.1, out_bf16=outb, colsum=cs), "relu": dict(bias=bias, relu=True, dropout_p=0.1, seed=3, out_bf16=outb),
          "resid_drop": dict(bias=bias, dropout_p=0.1, seed=3, resid=resid, out_f32=out), "resid": dict(resid=resid, out_f32=out),
          "bf16": dict(out_bf16=outb), "bias_bf16": dict(bias=bias, out_bf16=outb), "bias_f32": dict(bias=bias, out_f32=out),
          "acc_resid": dict(resid=resid, out_f32=out, accumulate=True)}[epi]
    fl = 2.0 * m * n * k
    base = timeit(lambda: ops.gemm(a, w, b_layout=bl, **kw))
    line = "%-26s %5dx%5dx%5d %-10s auto %5.1f us %5.0f TF/s |" % (name, m, n, k, epi, base * 1e6, fl / base / 1e12)
    best = (base, "auto")
    splits = (0,) if "accumulate" not in kw else (0, 2, 4, 8, 16)
    for cg, bn in ((2, 256), (2, 128), (1, 256), (1, 128), (1, 64)):
        if cg == 2 and m <= 128:
            continue
        for sk in splits:
            try:
                t = timeit(lambda: ops.gemm(a, w, b_layout=bl, cta_group=cg, block_n=bn, split_k=sk, **kw))
            except Exception as e:  # noqa: BLE001
                continue
            line += " cg%d/%d%s %5.1f" % (cg, bn, ("/s%d" % sk) if sk else "", t * 1e6)
            if t < best[0]:
                best = (t, "cg%d bn%d split%d" % (cg, bn, sk))
    print(line + " || best %s %5.1f us (%+.1f%%)" % (best[1], best[0] * 1e6, 100 * (best[0] / base - 1)), flush=True)
