"""One eager MCAN training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
    python tools/profile_step.py [large|small] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcan_vqa_b200.train import Trainer  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = Trainer(bench.Cfg(bench.MODELS[model]), bench.TOKEN_SIZE, bench.ANSWER_SIZE, dev, use_graph=False)
batch = bench.synth_batch(bench.BATCH, 1234, device=dev)
for _ in range(3):
    tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled %d step(s) of MCAN-%s" % (steps, model))
