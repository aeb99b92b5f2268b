"""fp32 + residual epilogue GEMMs (merge / FFN2 forward, their dgrads): auto-picked tile vs forced alternatives."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=15):
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


for m, n, k in ((6400, 1024, 1024), (6400, 1024, 4096), (6400, 1024, 3072)):
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda")
    out = torch.empty(m, n, device="cuda")
    outb = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * m * n * k
    for name, kw in (("bf16 out", dict(out_bf16=outb)), ("fp32 out + resid", dict(out_f32=out, resid=resid)),
                     ("fp32 + resid + bias + dropout", dict(out_f32=out, resid=resid, bias=bias, dropout_p=0.1, seed=5))):
        line = "%dx%dx%d %-30s" % (m, n, k, name)
        for cg, bn in ((0, 0), (2, 256), (2, 128), (1, 256), (1, 128)):
            t = timeit(lambda: ops.gemm(a, w, cta_group=cg, block_n=bn, **kw))
            line += " | cg%d bn%-3d %5.1f us %6.0f TF/s" % (cg, bn, t * 1e6, fl / t / 1e12)
        print(line, flush=True)
