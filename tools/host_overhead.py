"""Host-side cost of one eager MCAN training step with the C ABI stubbed out (no GPU needed): how long Python, ctypes
argument marshalling and the tensor allocator take for the ~300 launches of a step -- the part that bounds the eager
route the unchanged core/exec.py takes (with the CUDA graph of mcan_vqa_b200.train.Trainer it is paid once).

    python tools/host_overhead.py [large|small] [--profile]
Tensors live on the CPU here (malloc instead of the CUDA caching allocator), so the absolute figure is a lower bound of the
GPU-side host time; the ranking of the hot spots is what it is for."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
from mcan_vqa_b200 import blocks, capi, ops  # noqa: E402


class StubLib(object):
    def __init__(self):
        self.calls = 0

    def __getattr__(self, name):
        if not name.startswith("mcan_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls += 1
            return 148 if name == "mcan_num_sms" else 0
        setattr(self, name, fn)
        return fn


def main():
    model = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "large"
    stub = StubLib()
    capi.load = lambda: stub
    ops._req = lambda *a, **k: None
    ops._stream = lambda: 0
    ops._raw_sms = [0]
    blocks.DRY_RUN = True
    torch.zeros_like_real, torch.zeros_real = torch.zeros_like, torch.zeros
    torch.zeros = lambda *a, **k: torch.empty(*a, **k)          # no CPU memsets of the arenas: host logic only
    from core.model.net import Net
    cfg = bench.Cfg(bench.MODELS[model])
    net = Net(cfg, None, bench.TOKEN_SIZE, bench.ANSWER_SIZE).train()
    img, ques, ans = bench.synth_batch(bench.BATCH, 1234, device="cpu")

    def step():
        net.zero_grad(set_to_none=True)
        loss = net.forward_with_loss(img, ques, ans)[0]
        loss.backward()

    for _ in range(3):
        step()
    c0 = stub.calls
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    per = (stub.calls - c0) / n
    print("MCAN-%s: %.2f ms host time per forward+backward, %d C-ABI calls -> %.1f us per call" % (model, dt * 1e3, per, dt * 1e6 / per))
    if "--profile" in sys.argv:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            step()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
