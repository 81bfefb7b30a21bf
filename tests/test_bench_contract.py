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


@pytest.mark.parametrize("cfg,unit", [("C2", "env-steps/s"), ("C4", "env-steps/s"), ("C5", "solves/s")])
def test_reference_arm_covers_every_config(cfg, unit):
    j = _run(["--impl", "reference", "--config", cfg, "--steps", "1", "--warmup", "1", "--cpu-envs", "64"], 300)
    assert j["impl"] == "reference" and BASE_KEYS <= set(j) and j["unit"] == unit and j["value"] > 0
    assert j["config"]["config"] == cfg and "bounded sample" in j["config"]["reference_arm"] and j["ms_per_step_is_for_units"] == 64


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["--config", "C4", "--scaling", "weak", "--envs-per-gpu", "1024", "--neighbor-dist", "3"],
                                  ["--config", "C5", "--scaling", "weak", "--envs-per-gpu", "4096", "--horizon", "20"]])
def test_gpu_arm_other_configs_carry_a_clean_parity_gate(args):
    j = _run(args + ["--steps", "3", "--warmup", "3", "--cpu-envs", "128"], 900)
    assert BASE_KEYS | {"roofline", "roofline_compute", "clocks", "gpu_launches", "parity"} <= set(j)
    p = j["parity"]
    if "orca_velocity_mismatches" in p:
        assert p["orca_velocity_mismatches"] == 0 and p["lp_branch_divergences"] == 0 and p["max_position_err_m"] <= 1e-12
    else:
        assert p["lm_branch_divergences"] <= 3 and p["max_control_err_same_branch"] <= 5e-6
    assert j["roofline_compute"]["flops_source"].startswith(("instrumented", "survey formula on the in-range"))


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
    p = j["parity"]
    assert p["orca_velocity_mismatches"] == 0 and p["done_word_mismatches"] == 0 and p["lm_branch_divergences"] <= 2
    assert p["max_position_err_m"] <= 1e-12 and p["max_reward_err"] <= 1e-9 and p["max_obs_err"] <= 1e-9
    assert j["roofline_compute"]["flops_source"] == "instrumented" and j["lookahead_prune_variant"]["value"] > j["value"]
