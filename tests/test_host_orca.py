"""CPU: PRODUCT code of the hot kernel on the host. dr-mpc_b200/csrc/orca_device.cuh is written against a lane-group type; with
the one-lane SerialGroup (what the thread-per-agent CUDA kernel uses) the very same source -- neighbour ranking, half-plane
construction, the single LP2 / LP3 loop, linearProgram1, both neighbour-search variants (full ranking, cut-off + compaction) --
compiles for the host (tests/host_harness/orca_host.cpp, -ffp-contract=off like the CUDA unit's -fmad=false) and must give the
reference's golden velocities BIT FOR BIT and the oracle's decision counters field for field. The lane-parallel form of the
same code is what tests/test_gpu_orca.py runs on the B200."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from conftest import ORCA_KEYS, REPO


@pytest.fixture(scope="module")
def host_orca(oracle):
    import dr_mpc_b200._lib as L
    hh = os.path.join(REPO, "tests", "host_harness")
    so = os.path.join(hh, "liborca_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-pthread", "-Wno-unknown-pragmas",
                           "-I/usr/local/cuda/include", "-o", so, os.path.join(hh, "orca_host.cpp")])
    lib = ctypes.CDLL(so)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    def run(hpos, hvel, goal, rpos, neighbor_dist=10.0, cutoff=False, is_static=None, robot_visible=True, lanes=1):
        prm = L.OrcaParams(neighbor_dist, 5.0, 0.25, 0.3 + 0.01 + 0.175, 1.0, 0, int(robot_visible), 0)   # orca.py:79-89
        hpos = np.ascontiguousarray(hpos, dtype=np.float64)
        hvel = np.ascontiguousarray(hvel, dtype=np.float32)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        rpos = np.ascontiguousarray(rpos, dtype=np.float64)
        st8 = np.ascontiguousarray(is_static, dtype=np.uint8) if is_static is not None else None
        E, Nh, _ = hpos.shape
        v = np.zeros((E, Nh, 2), np.float32)
        st = np.zeros((E, Nh), dtype=oracle.ORCA_STATS_DTYPE)
        assert lib.host_orca_crowd_pass_lanes(ctypes.byref(prm), E, Nh, p(hpos), p(hvel), p(goal), p(rpos), p(st8), int(cutoff), int(lanes), p(v), p(st)) == 0
        return v, st
    return run


@pytest.mark.parametrize("cutoff", [False, True])
@pytest.mark.parametrize("key", ORCA_KEYS)
def test_product_orca_source_matches_reference_golden_bit_exact(host_orca, oracle, golden_orca, key, cutoff):
    g = golden_orca
    args = (g[key + "_hpos"], g[key + "_hvel"], g[key + "_goal"], g[key + "_rpos"])
    v, st = host_orca(*args, cutoff=cutoff)
    assert np.array_equal(v.view(np.uint32), g[key + "_vnew"].view(np.uint32)), np.abs(v - g[key + "_vnew"]).max()
    _, so = oracle.orca_crowd_pass(*args, want_stats=True)
    for f in ("n_lines", "lp2_fail", "n_lp1", "lp1_work", "branch_mask", "n_lp3_proj"):     # zero LP-branch divergences
        assert np.array_equal(st[f], so[f]), f


@pytest.mark.parametrize("Nh,spread,neighbor_dist", [(6, 3.0, 10.0), (20, 5.0, 10.0), (50, 8.0, 10.0), (50, 8.0, 3.0), (20, 5.0, 3.0), (64, 9.0, 2.0)])
def test_product_orca_source_matches_oracle_on_random_crowds(host_orca, oracle, Nh, spread, neighbor_dist):
    """Dense random crowds (up to the engine's maximum of 64 humans), the reference's 10 m range and BASELINE config 4's 3 m
    range, static humans and an invisible robot mixed in: velocities bit-exact, every decision counter equal, for both
    neighbour-search variants (they must agree with each other and with RVO2's kd-tree order)."""
    rng = np.random.default_rng(Nh * 7 + int(neighbor_dist))
    E = 250
    hpos = rng.uniform(-spread, spread, (E, Nh, 2))
    goal = rng.uniform(-spread, spread, (E, Nh, 2))
    hvel = rng.uniform(-0.7, 0.7, (E, Nh, 2)).astype(np.float32)
    rpos = rng.uniform(-spread, spread, (E, 2))
    static = (rng.uniform(size=(E, Nh)) < 0.1).astype(np.uint8)
    for visible in (True, False):
        vo, so = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, is_static=static, robot_visible=visible, neighbor_dist=neighbor_dist, want_stats=True)
        assert (so["lp2_fail"] < so["n_lines"]).any() or neighbor_dist < 10           # LP3 is exercised in the dense sets
        for cutoff in (False, True):
            v, st = host_orca(hpos, hvel, goal, rpos, neighbor_dist=neighbor_dist, cutoff=cutoff, is_static=static, robot_visible=visible)
            assert np.array_equal(v.view(np.uint32), vo.view(np.uint32)), (visible, cutoff, np.abs(v - vo).max())
            for f in ("n_lines", "lp2_fail", "n_lp1", "lp1_work", "branch_mask", "n_lp3_proj"):
                assert np.array_equal(st[f], so[f]), (visible, cutoff, f)


@pytest.mark.parametrize("lanes", [2, 4, 8, 16])
def test_lane_parallel_orca_source_on_emulated_lanes(host_orca, oracle, golden_orca, lanes):
    """The COOPERATIVE form of the same source -- what the 4- / 8- / 16-lane CUDA groups of ha_step_kernel and orca_predict_kernel
    execute: `lanes` OS threads play the lanes of one group, every shuffle / ballot / butterfly reduction / group sync of the
    group interface is an exchange between two barriers (tests/host_harness/orca_host.cpp: HostLaneGroup). Per-lane distance
    tests and rankings, lines stored at their sorted rank, lane-parallel violated-line scans and linearProgram1 min / max
    reductions must reproduce every golden crowd BIT FOR BIT with the oracle's decision counters, for both neighbour-search
    variants, and a short-range random crowd with static humans."""
    if lanes > 2 * (os.cpu_count() or 1):
        pytest.skip("spinning lane barriers need about as many cores as lanes")
    g = golden_orca
    for key in ORCA_KEYS:
        args = (g[key + "_hpos"], g[key + "_hvel"], g[key + "_goal"], g[key + "_rpos"])
        _, so = oracle.orca_crowd_pass(*args, want_stats=True)
        for cutoff in (False, True):
            v, st = host_orca(*args, cutoff=cutoff, lanes=lanes)
            assert np.array_equal(v.view(np.uint32), g[key + "_vnew"].view(np.uint32)), (key, cutoff)
            for f in ("n_lines", "lp2_fail", "n_lp1", "lp1_work", "branch_mask", "n_lp3_proj"):
                assert np.array_equal(st[f], so[f]), (key, cutoff, f)
    rng = np.random.default_rng(lanes)
    E, Nh, spread = 60, 50, 8.0
    hpos, goal = rng.uniform(-spread, spread, (E, Nh, 2)), rng.uniform(-spread, spread, (E, Nh, 2))
    hvel, rpos = rng.uniform(-0.7, 0.7, (E, Nh, 2)).astype(np.float32), rng.uniform(-spread, spread, (E, 2))
    static = (rng.uniform(size=(E, Nh)) < 0.1).astype(np.uint8)
    for nd in (3.0, 10.0):
        vo, so = oracle.orca_crowd_pass(hpos, hvel, goal, rpos, is_static=static, neighbor_dist=nd, want_stats=True)
        for cutoff in (False, True):
            v, st = host_orca(hpos, hvel, goal, rpos, neighbor_dist=nd, cutoff=cutoff, is_static=static, lanes=lanes)
            assert np.array_equal(v.view(np.uint32), vo.view(np.uint32)), (nd, cutoff)
            assert np.array_equal(st["lp1_work"], so["lp1_work"]) and np.array_equal(st["n_lp3_proj"], so["n_lp3_proj"])
