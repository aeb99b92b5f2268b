"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port on the host cores)
prints one JSON line with the keys the driver reads, and the FLOP model matches SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "small",
                        "--steps", "1", "--warmup", "3", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MCAN train samples/sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "MCAN-small" in d["config"]["workload"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_flop_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    small = bench.hot_path_train_flops_per_sample(bench.Cfg(bench.MODELS["small"]))
    large = bench.hot_path_train_flops_per_sample(bench.Cfg(bench.MODELS["large"]))
    assert small == 3 * (5163073536 + 61050880)          # MCA_ED + AttFlat, forward, x3 for training
    assert large == 3 * (20367310848 + 128276480)
