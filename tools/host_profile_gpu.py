"""cProfile of eager MCAN-large training steps on the GPU (host side of the route the unchanged core/exec.py takes):
    python tools/host_profile_gpu.py [steps]
Prints the wall time per eager step (host-bound) and the hottest Python functions."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcan_vqa_b200.train import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = Trainer(bench.Cfg(bench.MODELS["large"]), bench.TOKEN_SIZE, bench.ANSWER_SIZE, dev, use_graph=False)
batch = bench.synth_batch(bench.BATCH, 1234, device=dev)
for _ in range(5):
    tr.step(*batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    tr.step(*batch)
torch.cuda.synchronize()
print("eager step: %.2f ms wall" % ((time.perf_counter() - t0) / steps * 1e3))
t0 = time.perf_counter()
for _ in range(steps):
    tr.step(*batch)
host = (time.perf_counter() - t0) / steps * 1e3
torch.cuda.synchronize()
print("eager step: %.2f ms host time until the last launch is queued" % host)
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    tr.step(*batch)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumtime").print_stats(35)
print(s.getvalue())
