"""The bench.py contract on CPU: the reference arm prints one JSON line with the keys the driver reads (run here on the toy
workload with a small sample), ranks other than 0 print nothing, and the B200 arm refuses to run without a CUDA device (no
CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT, needs_ref

BENCH = os.path.join(ROOT, "bench.py")


@needs_ref
def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "toy", "--steps", "1", "--warmup", "0", "--cpu-sample", "2000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("reads/sec (krepp dist") and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["value"] > 100 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]
    # under torchrun only rank 0 runs and prints the reference arm
    r1 = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "toy", "--steps", "1", "--warmup", "0", "--cpu-sample", "2000"],
                        capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # (on a GPU box the arm runs; the -m gpu suite and the round scripts cover it)
    r = subprocess.run([sys.executable, BENCH, "--workload", "toy", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
