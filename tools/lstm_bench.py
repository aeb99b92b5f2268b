"""Question encoder (embedding + LSTM, forward and backward): library kernels vs torch (cuDNN, TF32) on the same GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from mcan_vqa_b200 import blocks, ops  # noqa: E402
from mcan_vqa_b200.blocks import LinearParams, Runtime  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / iters


for H in (1024, 512):
    B, T, E, V = 64, 14, 300, 20000
    emb = torch.nn.Embedding(V, E).cuda()
    lstm = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).cuda()
    tokens = torch.randint(1, V, (B, T), device="cuda")
    lp_ih = LinearParams([(lstm.weight_ih_l0, lstm.bias_ih_l0)], pad=64).get(True, True)
    lp_hh = LinearParams([(lstm.weight_hh_l0, lstm.bias_hh_l0)]).get(True)
    dq = torch.randn(B * T, H, device="cuda") * 0.1
    state = {}

    def fwd():
        rt = Runtime(True, 0.0)
        state["rt"] = rt
        state["out"] = blocks.qenc_fwd(rt, emb.weight.detach(), lp_ih, lp_hh, tokens, True)

    def bwd():
        blocks.qenc_bwd(state["rt"], state["out"][2], dq, V)

    def only_lstm_fwd():
        c = state["out"][2]
        xw = state["xw"]
        ops.lstm_fwd(xw, lp_hh.w, lp_hh.b, c.hbuf, state["out"][0], c.cbuf, c.gates, batch=B, steps=T, hidden=H)

    def only_lstm_bwd():
        c = state["out"][2]
        ops.lstm_bwd(dq, lp_hh.w, c.hbuf, c.cbuf, c.gates, state["da"], batch=B, steps=T, hidden=H)

    fwd()
    state["xw"] = torch.randn(B * (T + 1), 4 * H, device="cuda") * 0.1
    state["da"] = torch.empty(B * (T + 1), 4 * H, device="cuda", dtype=torch.bfloat16)
    tf, tb = timeit(fwd), timeit(bwd)
    kf, kb = timeit(only_lstm_fwd), timeit(only_lstm_bwd)

    def torch_fwd_bwd():
        emb.zero_grad(set_to_none=True)
        lstm.zero_grad(set_to_none=True)
        out, _ = lstm(emb(tokens))
        out.backward(dq.view(B, T, H))

    def torch_fwd():
        with torch.no_grad():
            lstm(emb(tokens))

    tt, ttf = timeit(torch_fwd_bwd), timeit(torch_fwd)
    print("H=%4d  library: fwd chain %6.1f us (LSTM kernel %6.1f us = %.1f us/step), bwd chain %6.1f us (LSTM kernel %6.1f us = %.1f us/step)"
          "  |  torch/cuDNN: fwd %6.1f us, fwd+bwd %6.1f us" % (H, tf * 1e6, kf * 1e6, kf * 1e6 / T, tb * 1e6, kb * 1e6, kb * 1e6 / T,
                                                                 ttf * 1e6, tt * 1e6), flush=True)
