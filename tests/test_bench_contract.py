"""bench.py's driver contract: exactly one JSON line on stdout with the agreed keys. The reference arm runs on the CPU
(oracle port) and is checked here; the GPU arm is checked with a short run on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--envs-per-gpu", "64", "--cpu-envs", "64"], 300)
    assert j["impl"] == "reference" and BASE_KEYS <= set(j)
    assert j["value"] > 0 and j["unit"] == "env-steps/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line_with_roofline_and_clocks():
    j = _run(["--steps", "6", "--warmup", "3", "--envs-per-gpu", "2048", "--humans", "10", "--cpu-envs", "256"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "kernels_ms"} <= set(j) and "impl" not in j
    assert j["n_gpus"] == 1 and j["steps"] == 6 and j["warmup"] >= 3 and j["value"] > 0 and j["gpu_launches"] > 0
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert j["e2e"]["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["e2e"]["value"] != j["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] > 0
