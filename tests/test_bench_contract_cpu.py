"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU -- it times the oracle port of the
reference's CPU path (the measured arm's full batch when it fits the time budget, else a bounded sample) -- and prints ONE
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    # a 25 s budget forces the bounded-sample branch on this container's 8 cores (the full 8+8 step takes ~90 s here)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-budget", "25"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ssl_train_step_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    gb = d["config"]["global_batch"]
    assert 2 <= gb <= 16 and d["config"]["same_batch_as_gpu_arm"] == (gb == 16)
    assert d["value"] > 0 and abs(d["value"] - gb / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
