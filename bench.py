#!/usr/bin/env python
"""MCAN training-throughput benchmark (BASELINE.json metric: MCAN train samples/sec at 1/2/4/8 B200).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # CPU arm: the unmodified reference (baseline/_ref) on the host cores

A step = one full training step of Net (embedding+LSTM, img_feat_linear, MCA_ED, AttFlat x2,
proj, sigmoid, BCE(sum), backward, AdamW) on one synthetic batch of 64 samples per GPU
(100 x 2048 region features, 14 tokens, 3129 answers), dropout 0.1, bf16 tensor-core GEMMs.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {
    "small": dict(hidden_size=512, multi_head=8, layer=6, flat_mlp_size=512, flat_glimpses=1, flat_out_size=512),
    "large": dict(hidden_size=1024, multi_head=16, layer=6, flat_mlp_size=512, flat_glimpses=1, flat_out_size=2048),
}
TOKEN_SIZE, ANSWER_SIZE, REGIONS, TOKENS, IMG_FEAT, BATCH = 20000, 3129, 100, 14, 2048, 64


class Cfg(object):
    def __init__(self, d, dropout_rate=0.1):
        self.__dict__.update(d)
        self.ff_size = 4 * self.hidden_size
        self.hidden_size_head = self.hidden_size // self.multi_head
        self.dropout_rate = dropout_rate
        self.word_embed_size = 300
        self.img_feat_size = IMG_FEAT
        self.use_glove = False


def hot_path_train_flops_per_sample(c):
    """3 x forward FLOPs of MCA_ED + 2 x AttFlat, dense padded shapes."""
    H, L, Sq, Sv, M, G, O = c.hidden_size, c.layer, TOKENS, REGIONS, c.flat_mlp_size, c.flat_glimpses, c.flat_out_size
    sa = 24 * Sq * H * H + 4 * Sq * Sq * H
    sga = (28 * Sv + 4 * Sq) * H * H + 4 * Sv * Sv * H + 4 * Sv * Sq * H
    flat = sum(2 * S * H * M + 2 * S * M * G + 2 * S * H * G + 2 * H * G * O for S in (Sq, Sv))
    return 3 * (L * (sa + sga) + flat)


def synth_batch(batch, seed, device=None, pin=False, ragged="none"):
    """SURVEY 8d synthetic inputs; ragged="prefix": n_v ~ U{10..100} regions and n_q ~ U{1..14} tokens per
    sample, the rest zero (BASELINE.json configs[3], the mask-heavy path)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, REGIONS, IMG_FEAT, generator=g).abs_()
    ques = torch.randint(1, TOKEN_SIZE, (batch, TOKENS), generator=g)
    if ragged == "prefix":
        n_v = torch.randint(10, REGIONS + 1, (batch,), generator=g)
        n_q = torch.randint(1, TOKENS + 1, (batch,), generator=g)
        img *= (torch.arange(REGIONS)[None, :] < n_v[:, None]).float()[:, :, None]
        ques *= (torch.arange(TOKENS)[None, :] < n_q[:, None]).long()
    ans = torch.zeros(batch, ANSWER_SIZE)
    idx = torch.randint(0, ANSWER_SIZE, (batch, 3), generator=g)
    val = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (batch, 3), generator=g)]
    ans.scatter_(1, idx, val)
    out = (img, ques, ans)
    if pin:
        out = tuple(t.pin_memory() for t in out)
    if device is not None:
        out = tuple(t.to(device) for t in out)
    return out


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def _host_threads():
    """All host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refload
    return refload.load()


def cpu_reference_step_time(model, batch, iters, warm, dropout=0.1, ragged="none"):
    """One full training step (Net forward, BCELoss(sum), backward, AdamW as core/model/optim.py builds it)
    of the UNMODIFIED reference (baseline/_ref, copied by oracle/fetch_ref.py) in fp32 on the host cores.
    Falls back to the oracle port (oracle/mcan_oracle.py, kind "port") when no copy of the reference is there.
    The only place bench.py executes anything under oracle/ or baseline/: the CPU baseline legs."""
    import torch
    threads = _host_threads()
    ref = _load_reference()
    cfg = Cfg(MODELS[model], dropout_rate=dropout)
    cfg.lr_base, cfg.batch_size = 1e-4 if model == "small" else 5e-5, batch
    img, ques, ans = synth_batch(batch, 1234, ragged=ragged)
    torch.manual_seed(0)
    times = []
    if ref is not None:
        kind = "reference"
        net = ref.net.Net(cfg, None, TOKEN_SIZE, ANSWER_SIZE).train()
        optim = ref.optim.get_optim(cfg, net, 64 * 1000)
        loss_fn = torch.nn.BCELoss(reduction="sum")
        for it in range(warm + iters):
            t0 = time.perf_counter()
            optim.zero_grad()
            loss = loss_fn(net(img, ques)[0], ans)
            loss.backward()
            optim.step()
            if it >= warm:
                times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        import mcan_oracle as orc
        ocfg = orc.Cfg(dropout_rate=dropout, **MODELS[model])
        ocfg.training = True
        sd = orc.synth_state_dict(ocfg, TOKEN_SIZE, ANSWER_SIZE, seed=0)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.AdamW(list(params.values()), lr=2.5e-5, weight_decay=1e-4)
        for it in range(warm + iters):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = orc.bce_sum(orc.net_forward(params, img, ques, ocfg)[0], ans)
            loss.backward()
            opt.step()
            if it >= warm:
                times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], threads, kind


def reference_gpu_eager(model, batch_dev, steps=10, warm=3):
    """The honest same-box bar: the unmodified reference Net, eager PyTorch on this B200,
    full training step with the reference's optimiser, fp32 (TF32 off = the reference's arithmetic) and under
    torch.autocast(bf16).  CUDA events.  None when no copy of the reference is available."""
    import torch
    ref = _load_reference()
    if ref is None:
        return None
    cfg = Cfg(MODELS[model])
    cfg.lr_base, cfg.batch_size = 1e-4 if model == "small" else 5e-5, BATCH
    out = {"implementation": "unmodified reference Net (baseline/_ref), eager PyTorch %s, AdamW via core/model/optim.py" % torch.__version__}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    loss_fn = torch.nn.BCELoss(reduction="sum")
    try:
        for name, tf32, autocast in (("fp32", False, False), ("tf32", True, False), ("autocast_bf16", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.manual_seed(0)
            net = ref.net.Net(cfg, None, TOKEN_SIZE, ANSWER_SIZE).cuda().train()
            optim = ref.optim.get_optim(cfg, net, 64 * 1000)
            img, ques, ans = batch_dev

            def step():
                optim.zero_grad()
                if autocast:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        probs = net(img, ques)[0]
                    loss = loss_fn(probs.float(), ans)      # BCELoss is not autocast-safe in bf16
                else:
                    loss = loss_fn(net(img, ques)[0], ans)
                loss.backward()
                optim.step()

            for _ in range(warm):
                step()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                step()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / steps
            out[name] = {"samples_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms}
            del net, optim
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    return out


def workload_name(model, ragged="none"):
    return ("MCAN-%s full training step (Net fwd + BCE(sum) + bwd + AdamW), batch %d per GPU, "
            "100x2048 region feats, 14 tokens, 3129 answers, dropout 0.1, random init%s" %
            (model, BATCH, "" if ragged == "none" else ", ragged 10-100 regions / 1-14 tokens"))


def make_config(model, world, ragged="none"):
    """`config` of the JSON line -- identical for the B200 arm and the reference arm."""
    return {"workload": workload_name(model, ragged), "global_batch": BATCH * world, "parallelism": "dp%d" % world,
            "l2": "working set per step (fp32 masters + bf16 copies + activations, > 1 GB) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the box's host cores
    (all threads), full 64-sample batches of the same workload; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = args.cpu_batch
    t, cores, kind = cpu_reference_step_time(args.model, sample_b, args.steps, args.warmup, ragged=args.ragged)
    val = sample_b / t
    sample = ("%d-sample training steps (fwd + BCE(sum) + bwd + AdamW, dropout 0.1) of the unmodified reference in fp32, "
              "median of %d timed after %d warm-up; %d threads on '%s'" %
              (sample_b, args.steps, args.warmup, cores, _cpu_model_name()))
    line = {
        "impl": "reference", "metric": "MCAN train samples/sec", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.model, args.gpus, args.ragged),
        "implementation": ("unmodified reference (baseline/_ref: core/model/net.py Net + core/model/optim.py get_optim), torch CPU fp32"
                           if kind == "reference" else "CPU oracle port of the reference (no copy of the reference on this box)"),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample,
                         "cpu_model": _cpu_model_name()},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_profile_values():
    """Counters of the dominant GEMM shape from the committed ncu capture (profiles/, named in the line): they are
    evidence from that capture, not measured in this run -- the live figure of this run is roofline.achieved."""
    name = "r02_ncu_full_gemm_6400x4096x1024.metrics.csv"
    path = os.path.join(ROOT, "profiles", name)
    vals = {}
    try:
        for line in open(path):
            parts = line.strip().split(",")
            if len(parts) == 3:
                try:
                    vals[parts[0]] = (float(parts[2]), parts[1])
                except ValueError:
                    pass
    except OSError:
        return None, None, None
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    try:
        traffic = sum(vals[k][0] * scale[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tensor = vals["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]
    except KeyError:
        return None, None, "profiles/" + name
    return traffic, tensor, "profiles/" + name


def gemm_roofline(trainer, batch_dev, peaks):
    """One instrumented training step: CUDA events around every tcgen05 GEMM launch (on the launching
    stream).  achieved = sum of algorithmic FLOPs / sum of launch durations.  A 40 ms single-CTA spin
    kernel is queued first so that the host runs ahead of the device: otherwise an eager step is
    launch-bound and every event pair would also time the host gap in front of its kernel."""
    import torch
    from mcan_vqa_b200 import capi, ops
    lib = capi.load()
    records = []
    real = ops.gemm

    def timed(a, b, **kw):
        a0 = a[0] if isinstance(a, (list, tuple)) else a
        b0 = b[0] if isinstance(b, (list, tuple)) else b
        nseg = len(a) if isinstance(a, (list, tuple)) else 1
        m, k = (a0.shape[0], a0.shape[1]) if kw.get("a_layout", 0) == 0 else (a0.shape[1], a0.shape[0])
        n = b0.shape[0] if kw.get("b_layout", 0) == 0 else b0.shape[1]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real(a, b, **kw)
        e.record()
        records.append((2.0 * m * n * k * nseg, s, e, (m, n, k, kw.get("a_layout", 0), kw.get("b_layout", 0),
                                                     sorted(x for x in kw if kw[x] is not None and x not in ("a_layout", "b_layout")))))

    real_grouped = ops.gemm_grouped

    def timed_grouped(problems, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_grouped(problems, **kw)
        e.record()
        k = problems[0][0].shape[0]
        records.append((sum(2.0 * a.shape[1] * b.shape[1] * k for a, b, _ in problems), s, e,
                        ("grouped", [(a.shape[1], b.shape[1]) for a, b, _ in problems], k, 1, 1, ["accumulate", "out_f32"])))

    # The single-GPU step overlaps the AdamW update of finished layers with the encoder half of the backward pass
    # (optim.EarlyStep): GEMMs that share the SMs and the HBM with that update take several times longer, which is
    # the point of the overlap but says nothing about the kernel.  The per-launch figures are therefore taken in a
    # step WITHOUT the overlap (every GEMM owns the GPU, as under ncu); `value` / `ms_per_step` keep the overlap.
    from mcan_vqa_b200 import optim as _optim
    saved_early = (trainer.early, _optim._early)
    trainer.early = None
    _optim.set_early(None)
    ops.gemm = timed
    ops.gemm_grouped = timed_grouped
    try:
        for _ in range(2):
            records.clear()
            capi.check(lib.mcan_debug_hog(1, int(0.040 * 1.9e9), 1024, torch.cuda.current_stream().cuda_stream), "hog")
            trainer._raw_step(*batch_dev)
            torch.cuda.synchronize()
    finally:
        ops.gemm = real
        ops.gemm_grouped = real_grouped
        trainer.early = saved_early[0]
        _optim.set_early(saved_early[1])
    flops = sum(r[0] for r in records)
    secs = sum(r[1].elapsed_time(r[2]) for r in records) * 1e-3
    if os.environ.get("MCAN_BENCH_DUMP"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", os.environ["MCAN_BENCH_DUMP"]), "w") as f:
            json.dump([{"shape": r[3], "us": r[1].elapsed_time(r[2]) * 1e3, "flops": r[0]} for r in records], f)
    achieved = flops / secs / 1e12
    burst = peaks.get("bf16_tflops") or 1590.0
    sustained = peaks.get("bf16_tflops_sustained") or 1400.0
    traffic, tensor_pct, profile = ncu_profile_values()
    return {"bound": "tensor", "achieved": achieved, "unit": "TFLOP/s",
            "frac_burst": achieved / burst, "frac_sustained": achieved / sustained,
            "peak_burst": burst, "peak_sustained": sustained,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst) / bf16_tflops_sustained" if peaks else
                           "fallback 1590 / 1400 (B200_PROFILING.md)",
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant shape (FFN1 forward
            # 6400x4096x1024: A + W read exactly once, the bf16 output stays in the 126 MB L2; algorithmic bytes
            # 73.9 MB) and the tensor-pipe counter of the same capture: parsed from the committed ncu summary named
            # in `ncu_profile` -- evidence of that capture, not of this run
            "traffic": traffic, "traffic_launch": "gemm_tcgen05_kernel<256,0,0,2,1> 6400x4096x1024 (one launch)",
            "tensor_pipe_active_pct_ncu": tensor_pct, "ncu_profile": profile,
            "kernel": "gemm_tcgen05_kernel (all %d launches of one training step; instrumented step without the optimiser overlap)" % len(records),
            "gemm_ms_per_step": secs * 1e3, "gemm_flops_per_step": flops}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="large", choices=sorted(MODELS))
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--cpu-batch", type=int, default=BATCH, help="samples per CPU-baseline step (default: the full batch)")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the CPU baseline and the reference-on-GPU legs")
    ap.add_argument("--ragged", default="none", choices=["none", "prefix"],
                    help="prefix: 10-100 valid regions and 1-14 valid tokens per sample (BASELINE.json configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the MCAN hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Data parallel: the gradient all-reduce overlaps the backward pass and takes a few SMs away from
    # the persistent GEMMs.  Measured at 8 x B200 (MCAN-large): static tile schedule 10.92 ms/step,
    # dynamic schedule (MCAN_GEMM_DYNAMIC=1: an occupied SM just claims fewer tiles) 11.05 ms,
    # NCCL capped at 16 / 8 CTAs 11.34 / 12.49 ms -- the exchange is bandwidth-, not SM-limited.
    # MCAN_DP_RESERVE_SMS=n: static schedule on #SMs - n, NCCL capped at n CTAs.
    reserve = int(os.environ.get("MCAN_DP_RESERVE_SMS", "0"))
    dynamic = world > 1 and reserve == 0 and os.environ.get("MCAN_GEMM_DYNAMIC", "0") != "0"
    if world > 1:
        if reserve > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(reserve))
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    __graft_entry__.build()
    from mcan_vqa_b200 import blocks, capi, ops
    from mcan_vqa_b200.train import Trainer
    if world > 1 and reserve > 0:
        ops.set_sm_limit(ops.num_sms() - reserve)
    ops.set_gemm_schedule(dynamic)

    cfg = Cfg(MODELS[args.model])
    torch.manual_seed(0)           # identical random-init replicas on every rank
    use_graph = not args.no_graph
    graph_note = "cuda-graph"
    trainer = Trainer(cfg, TOKEN_SIZE, ANSWER_SIZE, dev, lr_base=1e-4 if args.model == "small" else 5e-5,
                      data_size=64 * 1000 * world, batch_size=BATCH * world, use_graph=use_graph,
                      data_parallel=world > 1)
    host = synth_batch(BATCH, 1234 + rank, pin=True, ragged=args.ragged)
    batch_dev = tuple(t.to(dev) for t in host)
    if use_graph:
        try:
            trainer.capture(*batch_dev)
        except Exception as e:  # capture is an optimisation; fall back to eager launches and say so
            graph_note = "eager (graph capture failed: %s)" % str(e).split("\n")[0][:120]
            trainer.graph = None
            trainer.use_graph = False
            torch.cuda.synchronize()
    else:
        graph_note = "eager"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident throughput --------------------------------------------------------
    for _ in range(args.warmup):
        trainer.step(*batch_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    secs = timed_region(lambda i: trainer.step(*batch_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(trainer.loss.item()) if trainer.graph is not None else None
    value = world * BATCH * args.steps / secs

    # ---- end to end: pinned host -> device copies and a device -> host read every step ------
    copy_stream = torch.cuda.Stream()
    bufs = [tuple(torch.empty_like(t, device=dev) for t in host) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = torch.zeros(args.steps + args.warmup, dtype=torch.float32).pin_memory()

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            for d, h in zip(bufs[slot], host):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    state = {"i": 0}
    prefetch(0)
    prefetch(1)

    # simple, correct double buffering: slot i&1 is refilled right after step i consumed it
    def e2e_step2(_):
        i = state["i"]
        slot = i & 1
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = trainer.step(*bufs[slot])
        done = torch.cuda.Event()
        done.record()
        copy_stream.wait_event(done)
        prefetch(slot)
        loss_host[i % loss_host.numel()].copy_(loss, non_blocking=True)
        state["i"] = i + 1

    for _ in range(args.warmup):
        e2e_step2(0)
    e2e_secs = timed_region(e2e_step2, args.steps)
    e2e_value = world * BATCH * args.steps / e2e_secs
    h2d = sum(t.numel() * t.element_size() for t in host)

    # ---- instrumented eager steps (all ranks take part: the backward all-reduces) -------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    probe = Trainer.__new__(Trainer)
    probe.__dict__.update(trainer.__dict__)
    probe.graph, probe.use_graph = None, False
    c0 = capi.launch_count
    probe._raw_step(*batch_dev)      # kernels per step, counted on an eager step (the graph replays exactly these)
    torch.cuda.synchronize()
    per_step = capi.launch_count - c0
    roof = gemm_roofline(probe, batch_dev, peaks)
    # the eager route: the same kernels launched one by one from Python, as the reference's unchanged core/exec.py
    # drives the overlay (no CUDA graph); host-bound, wall clock with a synchronize on both sides
    eager_ms = None
    if world == 1:
        for _ in range(2):
            probe._raw_step(*batch_dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            probe._raw_step(*batch_dev)
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) / 10 * 1e3

    # data parallel sanity: every replica applied the same summed gradients, so the parameters of all
    # ranks must still be bit-identical after ~100 optimiser steps (per-parameter checksums, max - min over ranks)
    divergence = None
    if world > 1:
        sums = torch.stack([p.detach().double().sum() for p in trainer.net.parameters()])
        hi, lo = sums.clone(), sums.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        divergence = float((hi - lo).abs().max().item())

    if rank == 0:
        flops_sample = hot_path_train_flops_per_sample(cfg)
        # the peak that fits this run: the timed region is a fraction of a second at the clock the sampler saw;
        # burst peak unless the clock sat well below its maximum (a long, power-capped run)
        at_burst = not clocks or not clocks.get("sm_mhz") or not clocks.get("sm_max_mhz") or \
            clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"]
        roof["peak"] = roof["peak_burst"] if at_burst else roof["peak_sustained"]
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["peak_choice"] = ("burst (sampled SM clock %s of %s MHz)" if at_burst else "sustained (sampled SM clock %s of %s MHz)") % (
            clocks.get("sm_mhz") if clocks else None, clocks.get("sm_max_mhz") if clocks else None)
        line = {
            "metric": "MCAN train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": make_config(args.model, world, args.ragged),
            "precision_mode": "bf16 GEMM/attention operands, fp32 accumulation / residual stream / LayerNorm / softmax "
                              "(training mode; the split-precision 'fp32' mode that reaches >= 99.9 % top-1 agreement is inference-only)",
            "details": {"launch": graph_note, "sms_reserved_for_nccl": reserve if world > 1 else 0,
                        "gemm_tile_schedule": "dynamic" if dynamic else "static",
                        "grad_exchange": ("all-reduce(SUM), %s, buckets >= %s MB" % (
                            (trainer.sync.compress or "fp32") + (" (fp32 gradients cast by the library, AdamW reads the bf16 sums in place)"
                                                                if trainer.sync.compress == "bf16" else ""),
                            os.environ.get("MCAN_DP_BUCKET_MB", "192"))) if world > 1 else "none",
                        "optimizer": "fused multi-tensor AdamW (library kernel, emits the bf16 operand copies)",
                        "attention": ("image self-attention (100 x 100): tcgen05 / TMEM kernels; question side and guided "
                                      "attention (<= 32 keys): mma.sync kernels") if os.environ.get("MCAN_ATTN_TC", "1") != "0"
                        else "mma.sync kernels (MCAN_ATTN_TC=0)",
                        "decoder_wgrads": "second stream, next to the encoder backward" if blocks.OVERLAP_WGRAD else "inline"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_secs / args.steps * 1e3},
            "gpu_launches": per_step * args.steps,
            "kernels_per_step": per_step,
            "roofline": roof,
            "step_mfu": {"hot_path_train_flops_per_sample": flops_sample,
                         "achieved_tflops": value / world * flops_sample / 1e12,
                         "frac_of_burst_peak": value / world * flops_sample / 1e12 / roof["peak_burst"],
                         "frac_of_sustained_peak": value / world * flops_sample / 1e12 / roof["peak_sustained"]},
            "loss": loss_val,
        }
        if divergence is not None:
            line["replica_checksum_divergence"] = divergence
        if eager_ms is not None:
            line["eager_route"] = {"ms_per_step": eager_ms, "value": BATCH / eager_ms * 1e3, "unit": "samples/s",
                                   "note": "same kernels launched one by one from Python, no CUDA graph (the route the "
                                           "unchanged core/exec.py takes); host-bound, 10 steps, wall clock"}
        if world == 1 and not args.skip_cpu:
            # the honest same-box bar: the unmodified reference, eager PyTorch on this GPU
            del trainer.graph
            line["reference_gpu_eager"] = reference_gpu_eager(args.model, batch_dev)
            # reported CPU baseline: the unmodified reference on the host cores, full 64-sample batches
            t, cores, kind = cpu_reference_step_time(args.model, args.cpu_batch, 2, 1, dropout=0.1, ragged=args.ragged)
            t0, _, _ = cpu_reference_step_time(args.model, args.cpu_batch, 1, 0, dropout=0.0, ragged=args.ragged)
            line["cpu_baseline"] = {
                "value": args.cpu_batch / t, "unit": "samples/s", "cores": cores, "kind": kind, "cpu_model": _cpu_model_name(),
                "sample": "%d-sample training steps (fwd + BCE(sum) + bwd + AdamW) of the %s in fp32, dropout 0.1, median of 2 "
                          "timed after 1 warm-up; dropout 0.0: 1 timed step" %
                          (args.cpu_batch, "unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port"),
                "value_dropout0": args.cpu_batch / t0}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Orderly teardown: release the captured graphs (they reference the communicator), drain the device,
        # then destroy the process group.  A watchdog exits the process if NCCL's teardown still blocks.
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        trainer.graph = None
        probe.graph = None
        trainer.close()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
