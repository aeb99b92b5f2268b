"""Times the AttFlat pool kernels (forward / backward) at the MCAN shapes, L2 flushed between launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
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


for name, B, S, H, M, G in (("image  large", 64, 100, 1024, 512, 1), ("question large", 64, 14, 1024, 512, 1),
                            ("image  small", 64, 100, 512, 512, 1)):
    hmid = torch.relu(torch.randn(B * S, M, device="cuda")).to(torch.bfloat16)
    w2 = torch.randn(G, M, device="cuda") * 0.1
    b2 = torch.randn(G, device="cuda")
    x = torch.randn(B * S, H, device="cuda")
    mask = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    att_w = torch.empty(B, S, G, device="cuda")
    p32 = torch.empty(B, G * H, device="cuda")
    pbf = torch.empty(B, G * H, device="cuda", dtype=torch.bfloat16)
    dpooled = torch.randn(B, G * H, device="cuda")
    dx = torch.empty(B * S, H, device="cuda")
    dh = torch.empty(B * S, M, device="cuda", dtype=torch.bfloat16)
    dw2 = torch.zeros(G, M, device="cuda")
    db2 = torch.zeros(G, device="cuda")
    kw = dict(batch=B, s=S, h=H, mlp=M, glimpses=G)
    tf = timeit(lambda: ops.attflat_pool_fwd(hmid, w2, b2, mask, x, att_w=att_w, pooled_f32=p32, pooled_bf16=pbf, **kw))
    tb = timeit(lambda: ops.attflat_pool_bwd(dpooled, p32, hmid, w2, mask, x, att_w, gate_scale=1.0, dx=dx, dhmid=dh,
                                             dw2=dw2, db2=db2, **kw))
    bf = B * S * (4.0 * H + 2 * M)
    bb = B * S * (8.0 * H + 4 * M)
    print("%s fwd %6.1f us %5.0f GB/s | bwd %6.1f us %5.0f GB/s" % (name, tf * 1e6, bf / tf / 1e9, tb * 1e6, bb / tb / 1e9), flush=True)
