"""CPU: the C ORCA oracle (oracle/rvo2_ref.c, brute-force neighbour order) against the golden vectors produced by the
reference's own crowd_sim.get_human_actions -> orca.py -> rvo2 shim (full simulator with RVO2's kd-tree)."""
import numpy as np
import pytest
from conftest import ORCA_KEYS


@pytest.mark.parametrize("key", ORCA_KEYS)
def test_oracle_crowd_pass_matches_reference_bit_exact(oracle, golden_orca, key):
    g = golden_orca
    v, st = oracle.orca_crowd_pass(g[key + "_hpos"], g[key + "_hvel"], g[key + "_goal"], g[key + "_rpos"], want_stats=True)
    ref = g[key + "_vnew"]
    # fp32 both sides, same operation order: bit-exact (kd-tree visit order only matters on exact distSq ties)
    assert np.array_equal(v.view(np.uint32), ref.view(np.uint32)), np.abs(v - ref).max()


def test_golden_covers_all_branches(oracle, golden_orca):
    g = golden_orca
    mask, lp3, nlp1 = 0, 0, 0
    for key in ORCA_KEYS:
        _, st = oracle.orca_crowd_pass(g[key + "_hpos"], g[key + "_hvel"], g[key + "_goal"], g[key + "_rpos"], want_stats=True)
        mask |= int(np.bitwise_or.reduce(st["branch_mask"].ravel()))
        lp3 += int((st["lp2_fail"] < st["n_lines"]).sum())
        nlp1 += int(st["n_lp1"].sum())
    assert mask == 0b1111          # cut-off circle, left leg, right leg, collision
    assert lp3 > 0 and nlp1 > 0    # infeasible LP2 -> LP3 exercised


def test_speed_limit_and_static(oracle):
    rng = np.random.default_rng(0)
    E, Nh = 64, 12
    hpos = rng.uniform(-4, 4, (E, Nh, 2)); goal = -hpos
    hvel = rng.uniform(-1, 1, (E, Nh, 2)).astype(np.float32)
    rpos = rng.uniform(-4, 4, (E, 2))
    static = rng.uniform(size=(E, Nh)) < 0.2
    v, st = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, is_static=static, want_stats=True)
    assert np.all(v[static] == 0)
    speed = np.linalg.norm(v.astype(np.float64), axis=-1)
    feasible = st["lp2_fail"] == st["n_lines"]
    assert np.all(speed[feasible] <= 1.0 + 1e-5)      # LP2 result lies in the max-speed disc
    assert np.mean(speed > 1.01) < 1e-2               # LP3 (infeasible) results may overshoot (see below)


def test_lp3_overshoot_is_reference_behaviour(oracle):
    """Two overlapping neighbours whose collision half-planes are almost anti-parallel: the projected line of
    linearProgram3 lies ~7e3 away, RVO2's fp32 discriminant dp^2 + r^2 - |p|^2 cancels and the 'clipped' velocity
    leaves the max-speed disc. This is what the restated RVO2 arithmetic does; the engine must reproduce it."""
    rng = np.random.default_rng(1)
    E, Nh = 2000, 50
    hpos = ((rng.uniform(size=(E, Nh, 2)) - 0.5) * 12).astype(np.float32).astype(np.float64)
    hvel = (rng.uniform(size=(E, Nh, 2)) - 0.5).astype(np.float32)
    rpos = (rng.uniform(size=(E, 2)) - 0.5) * 12
    v, st = oracle.orca_crowd_pass(hpos, hvel, -hpos, rpos, want_stats=True)
    speed = np.linalg.norm(v.astype(np.float64), axis=-1)
    over = speed > 1.01
    assert over.any() and (st["lp2_fail"][over] < st["n_lines"][over]).all() and over.mean() < 1e-3


def test_single_agent_moves_to_goal(oracle):
    v, st = oracle.orca_solve([0, 0, 0, 0, 0.6, 0.8, 0.485, 1.0], np.zeros((0, 5)))
    assert np.allclose(v, [0.6, 0.8]) and st["n_lines"] == 0


def test_instrumented_oracle_counts_and_agrees(oracle, golden_orca):
    """oracle/opcount: the flat oracle compiled UNCHANGED with instrumented scalar types (the source of bench.py's
    algorithmic flop figures, profiles/roofline.json). The instrumentation must not change a bit of the results, and the
    count per ORCA solve must sit where the survey's formula 53 n + sum(10 i + 20) puts it."""
    import ctypes
    import os
    import subprocess
    from conftest import REPO
    d = os.path.join(REPO, "oracle", "opcount")
    subprocess.check_call(["make", "-C", d, "-B", "libopcount.so"], stdout=subprocess.DEVNULL)
    cnt = ctypes.CDLL(os.path.join(d, "libopcount.so"))
    key = "n20_s5.0"
    hpos, hvel, goal, rpos, vnew = (golden_orca[f"{key}_{k}"] for k in ("hpos", "hvel", "goal", "rpos", "vnew"))
    fast = oracle.lib()
    try:
        oracle._lib = cnt
        cnt.opcount_reset()
        v, st = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, nthreads=1, want_stats=True)
        ops = (ctypes.c_uint64 * 8)()
        cnt.opcount_get(ops)
    finally:
        oracle._lib = fast
    assert np.array_equal(v.view(np.uint32), vnew.view(np.uint32))
    n = hpos.shape[0] * hpos.shape[1]
    flops = sum(list(ops)[:6]) / n
    formula = 53.0 * 20 + 10.0 * st["lp1_work"].mean() + 20.0 * st["n_lp1"].mean()
    print(f"counted {flops:.0f} flops per solve, survey formula {formula:.0f}")
    assert 0.7 * formula < flops < 1.5 * formula
    import json
    rj = json.load(open(os.path.join(REPO, "profiles", "roofline.json")))
    assert {"n6", "n10", "n20", "n50"} <= set(rj["orca"]) and {"K4", "K10", "K20", "K30", "K40"} <= set(rj["mpc"])
