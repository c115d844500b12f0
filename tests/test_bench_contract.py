"""CPU: the parts of the bench.py contract that do not need a GPU -- the reference arm prints ONE JSON line with the
agreed keys; the GPU arm refuses to run (exit code 1, no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spectra/sec" and d["unit"] == "spectra/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "sdss100k_predict"
    # the unmodified reference (baseline/_ref or /root/reference) when importable, else the dense port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["warmup"] == 1                        # the arm honours --steps / --warmup
    # same config keys as the GPU arm prints (the driver compares the two dicts)
    assert {"workload", "baseline_config", "kind", "grid", "Npix", "Nb", "Nh", "spectra_per_gpu_per_step", "precision",
            "l2"} <= set(d["config"])


def test_reference_arm_follows_the_headline_workload_of_the_gpu_arm():
    """N > 1: the headline is the data-parallel train step (configs[3]), so the reference arm times QFA.forward."""
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"]["workload"] == "sdss_train" and d["config"]["kind"] == "train" and d["n_gpus"] == 2
    assert d["config"]["allreduce"].startswith("none")          # same key set as the GPU arm's N > 1 config
    assert d["value"] > 0


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode == 1
    assert "no CUDA device" in json.loads(r.stdout.strip().splitlines()[-1])["error"]
