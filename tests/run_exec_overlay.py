"""Runs the reference's UNMODIFIED training loop (core/exec.py Execution.train) on the overlay.

    python tests/run_exec_overlay.py <workdir> <result.json>

sys.path = [this repository, baseline/_ref]: `core` is a namespace package in both trees, so
`core.exec`, `core.data.*`, `cfgs.*`, `utils.*` come from the reference and `core.model.*` from the
overlay (the drop-in boundary).  The dataset is synthetic (recipe: SURVEY 8c); everything
else -- Cfgs, CustomLoader, Net2 construction, .cuda(), get_optim, the step loop with per-parameter
gradient norms, the checkpoint -- is the reference's code, untouched.  Afterwards the checkpoint is
loaded into the reference's own Net2 and its probabilities are compared with the overlay's.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refload  # noqa: E402

workdir, result_path = sys.argv[1], sys.argv[2]
REF = refload.reference_root()
assert REF is not None, "no reference copy: run oracle/fetch_ref.py in the build container"
sys.path[:] = [ROOT, REF] + [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, REF)]
refload.stub_missing_modules()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

os.makedirs(os.path.join(workdir, "results"), exist_ok=True)
os.makedirs(os.path.join(workdir, "ck"), exist_ok=True)
os.chdir(workdir)

from cfgs.base_cfgs import Cfgs  # noqa: E402  (reference)
import core.exec as ref_exec  # noqa: E402       (reference)
import core.model.net as model_net  # noqa: E402 (overlay)
from mcan_vqa_b200 import capi  # noqa: E402

BATCH, STEPS_PER_EPOCH, P, S, TOKENS, ANSWERS = 64, 8, 100, 14, 2000, 3129

opt = Cfgs()
with open(os.path.join(REF, "cfgs", "small_model.yml")) as f:
    opt.add_args(yaml.safe_load(f))
opt.add_args({"run_mode": "train", "img_feat_pad_size": P, "use_glove": False, "gpu": "0", "seed": 7,
              "batch_size": BATCH, "max_epoch": 2, "eval_every_epoch": False, "resume": False, "num_workers": 0,
              "pin_mem": False, "verbose": True, "ckpt_path": os.path.join(workdir, "ck") + "/",
              "log_path": workdir + "/", "grad_norm_clip": -1})
opt.proc()


class SyntheticVQA(torch.utils.data.Dataset):
    """(img f32[P,2048], ques i64[14], ans f32[A], idx i64[1]) -- the collate contract of load_data.py:294-300."""

    def __init__(self, n):
        rs = np.random.RandomState(11)
        self.data_size, self.token_size, self.ans_size, self.pretrained_emb = n, TOKENS, ANSWERS, None
        self.img = np.abs(rs.standard_normal((n, P, opt.img_feat_size))).astype(np.float32)
        self.ques = rs.randint(1, TOKENS, size=(n, S)).astype(np.int64)
        self.ans = np.zeros((n, ANSWERS), np.float32)
        for i in range(n):
            self.img[i, rs.randint(10, P + 1):] = 0.0           # ragged region counts
            self.ques[i, rs.randint(1, S + 1):] = 0             # ragged question lengths
            # a learnable target: the answer is a function of the first token
            self.ans[i, int(self.ques[i, 0]) % ANSWERS] = 1.0

    def __len__(self):
        return self.data_size

    def __getitem__(self, idx):
        return (torch.from_numpy(self.img[idx]), torch.from_numpy(self.ques[idx]), torch.from_numpy(self.ans[idx]),
                torch.tensor([idx]))


dataset = SyntheticVQA(BATCH * STEPS_PER_EPOCH)
ex = ref_exec.Execution.__new__(ref_exec.Execution)      # __init__ only loads the real datasets
ex.opt, ex.model = opt, None
launches0 = capi.launch_count
model = ex.train(dataset)
torch.cuda.synchronize()
launches = capi.launch_count - launches0

log = open(os.path.join(workdir, "log_run_7.txt")).read()
epochs = re.findall(r"epoch = (\d+); loss = ([0-9.eE+-]+); lr = ([0-9.eE+-]+)", log)
ckpt_file = os.path.join(workdir, "ck", "ckpt_7", "epoch2.pt")
ckpt = torch.load(ckpt_file)

# the checkpoint in the reference's own Net2, fp32 eager on the GPU
ref = refload.load()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ref_net = ref.net.Net2(opt, None, TOKENS, ANSWERS)
missing = ref_net.load_state_dict(ckpt["state_dict"], strict=True)
ref_net.cuda().eval()
model.eval()
with torch.no_grad():
    v = torch.from_numpy(dataset.img[:BATCH]).cuda()
    q = torch.from_numpy(dataset.ques[:BATCH]).cuda()
    p_ref = ref_net(v, q)[0]
    p_new = model(v, q)[0]
rel = ((p_new - p_ref).abs() / p_ref.abs().clamp_min(1e-6)).max().item()
top1 = (p_new.argmax(1) == p_ref.argmax(1)).float().mean().item()

json.dump({
    "exec_file": os.path.abspath(ref_exec.__file__),
    "model_class": type(model).__module__ + "." + type(model).__name__,
    "model_file": os.path.abspath(model_net.__file__),
    "optimizer_keys": sorted(ckpt["optimizer"].keys()),
    "native_launches": launches,
    "epochs": len(epochs),
    "loss_epoch1": float(epochs[0][1]), "loss_epoch2": float(epochs[1][1]),
    "lr_epoch1": float(epochs[0][2]), "lr_epoch2": float(epochs[1][2]), "lr_base": float(opt.lr_base),
    "ckpt_probs_max_rel_vs_reference_net2": rel, "ckpt_top1_agreement": top1,
    "steps": 2 * STEPS_PER_EPOCH, "batch": BATCH,
}, open(result_path, "w"))
print("exec.py through the overlay: ok", launches, "kernel launches, probs rel err", rel)
