"""CPU: host-side logic and the C-ABI surface (no GPU compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
from conftest import REPO, mpc_path_table


def test_libdre_loads_and_exports_every_declared_symbol():
    import dr_mpc_b200._lib as L
    lib = L.load()
    hdr = open(os.path.join(REPO, "include", "dre.h")).read()
    names = sorted(set(re.findall(r"\b(dre_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libdre.so does not export {n}"
    assert lib.dre_version() >= 100
    assert lib.dre_strerror(-1).decode().startswith("invalid")


def test_struct_layouts_match_header_sizes():
    import dr_mpc_b200._lib as L
    assert ctypes.sizeof(L.OrcaParams) == 32 and ctypes.sizeof(L.OrcaStats) == 24
    assert ctypes.sizeof(L.MpcStatus) == 32 and ctypes.sizeof(L.MpcParams) == 72
    # compile a tiny C program against the header to cross-check the big struct
    src = '#include <stdio.h>\n#include "dre.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(dre_env_config), sizeof(dre_mpc_params), sizeof(dre_env_state_view), sizeof(dre_step_out));return 0;}'
    exe = os.path.join(REPO, "tests", "host_harness", "sizes")
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(REPO, "include"), "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    out = subprocess.check_output([exe]).decode().split()
    os.remove(exe)
    assert [int(x) for x in out] == [ctypes.sizeof(L.EnvConfig), ctypes.sizeof(L.MpcParams), ctypes.sizeof(L.EnvStateView), ctypes.sizeof(L.StepOut)]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "dr-mpc_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"oracle[/\\.](oracle|rvo2_ref|mpc_ref|env_ref|shims)", r"#include\s+\"[^\"]*oracle"):
                    assert not re.search(pat, txt, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import dr_mpc_b200._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.DreError, match="no CPU fallback"):
        L.load()


def test_paths_match_reference_paths_bit_exact(golden_mpc):
    """dr_mpc_b200.paths regenerates the reference's four continuous-task paths (path_tracking_env.py:28-56)."""
    import dr_mpc_b200 as d
    ours = d.paths.continuous_task_paths()
    for i, p in enumerate(ours):
        ref = golden_mpc[f"path{i}"]
        assert p.shape == ref.shape and np.abs(p - ref).max() <= 1e-13, i
    tab, lens = d.paths.pack_paths(ours)
    assert tab.shape == (4, 3, 125) and list(lens) == [125, 41, 125, 41]
    # reference quirk: path 2 (CCW from theta = pi) wraps to about -3.09 at node 1
    assert ours[2][2, 0] == np.pi and ours[2][2, 1] < -3.0


def test_engine_config_reads_reference_config_objects():
    import dr_mpc_b200 as d

    class B:
        pass

    def bag(**kw):
        b = B()
        b.__dict__.update(kw)
        return b
    cm = bag(
        config_general=bag(env=bag(lookback=6, time_step=0.25, time_limit=50), robot=bag(v_max=1.0, w_max=1.0)),
        config_HA=bag(sim=bag(human_num=7, circle_radius=4, circle_radius_humans=5), orca=bag(neighbor_dist=10, time_horizon=5, safety_space=0.175),
                      humans=bag(radius=0.3, v_max=1.0, end_goal_changing=True), robot=bag(radius=0.3, visible=True),
                      rewards=bag(params={'actuation_termination': {'min_vel': 0.025, 'penalty': -20}, 'safety_human_raise': {'safety_dist': 0.22, 'penalty': -15},
                                          'collision': {'penalty': -15}, 'disturbance': {'radius': 1.5, 'factor_v': 0.8 * 7, 'factor_th': 0.5 * 7}})),
        config_PT=bag(env=bag(node_separation=0.2), MPC=bag(planning_horizon=4, actions_in_input=4), PT_model_params={'MLP_local': bag(lookahead=20, lookbehind=5)},
                      rewards=bag(overall_scaling=0.5, params={'goal': {'goal_radius': 0.3, 'goal_angle': 0.5, 'goal_reward': 0}, 'path_advancement': {'path_advancement_scaling': 2.0},
                                                               'deviation': {'xy': 1, 'th': 0.05, 'scale': 0.9}, 'corridor_hit': {'penalty': -20, 'max_deviation': 1.6},
                                                               'deviation_end_of_path': {'penalty': -10}, 'safety_corridor_raise': {'penalty': -20, 'max_deviation': 1.4}})))
    c = d.EngineConfig.from_config_master(cm, num_envs=128)
    assert c.num_humans == 7 and c.num_envs == 128
    assert c == d.EngineConfig(num_humans=7, num_envs=128)             # the defaults ARE the reference's values
    e = c.env_config()
    assert abs(e.orca.radius - 0.485) < 1e-7 and e.mpc.horizon == 4 and e.safety_corridor_max_dev == 1.4
    assert e.num_static_humans == 0 and e.soft_reset_on_device == 0 and e.orca_dedup == 0
    # include_static_humans on (config_HA.py:46): the continuous task parks two obstacles after the dynamic humans
    cm.config_HA.sim.include_static_humans = {'episode_probability': 1, 'max_static_humans': 2}
    cm.config_general.env.continuous_task = True
    c2 = d.EngineConfig.from_config_master(cm, num_envs=128)
    e2 = c2.env_config()
    assert c2.num_humans == 9 and e2.num_static_humans == 2 and list(e2.static_pos)[:4] == [-4.0, 0.0, 0.0, 0.0]


def test_device_mpc_header_on_host_matches_oracle(oracle, golden_mpc):
    """The kernel's per-problem LM code (dr-mpc_b200/csrc/mpc_device.cuh), compiled for the host by
    tests/host_harness, against the oracle with teacher-forced warm starts."""
    hh = os.path.join(REPO, "tests", "host_harness")
    so = os.path.join(hh, "libmpc_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
                           "-o", so, os.path.join(hh, "mpc_host.cpp")])
    import dr_mpc_b200._lib as L
    lib = ctypes.CDLL(so)
    _, tab, lens = mpc_path_table(golden_mpc)
    for K, N in ((4, 600), (10, 200)):
        rng = np.random.default_rng(K)
        pid = rng.integers(0, 4, N).astype(np.int32)
        idx = rng.uniform(0, lens[pid] - 1)
        pose = np.stack([tab[pid, r, np.floor(idx).astype(int)] for r in range(3)], 1) + rng.uniform(-0.3, 0.3, (N, 3))
        ob = oracle.MpcBatch(N, oracle.mpc_params(K=K), tab, lens)
        prm = L.MpcParams(K, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)
        for step in range(2):
            wT, wu, it0 = ob.warm_T.copy(), ob.warm_u.copy(), ob.iteration0.astype(np.uint8)
            u_o, st_o = ob.act(pid, idx, pose)
            u_h = np.zeros((N, 2 * K)); st_h = np.zeros(N, dtype=oracle.MPC_STATS_DTYPE)
            p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            assert lib.host_mpc_act(ctypes.byref(prm), N, p(tab), 125, p(lens), p(pid), p(idx), p(pose), p(wT), p(wu), p(it0), p(u_h), p(st_h)) == 0
            same = (st_o["iterations"] == st_h["iterations"]) & (st_o["lm_solves"] == st_h["lm_solves"]) & (st_o["cost_evals"] == st_h["cost_evals"])
            assert same.mean() > 0.99 and np.abs(u_o - u_h)[same].max() < 1e-6
            assert np.abs(wT - ob.warm_T)[same].max() < 1e-5


def test_soft_reset_machine_on_host_follows_reference(golden_mpc):
    """The kernel's soft-reset decision code (dr-mpc_b200/csrc/env_device.cuh: soft_reset_decide), compiled for the host,
    replays every env.step call the reference's own HAAndPTEnv.soft_reset issued (tests/golden/soft_reset.npz): same
    scripted-or-caller decision and the same scripted action, bit for bit, at every call."""
    hh = os.path.join(REPO, "tests", "host_harness")
    so = os.path.join(hh, "libenv_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
                           "-o", so, os.path.join(hh, "env_host.cpp")])
    lib = ctypes.CDLL(so)
    g = np.load(os.path.join(REPO, "tests", "golden", "soft_reset.npz"))
    _, tab, lens = mpc_path_table(golden_mpc)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    seen = set()
    for tag in ("h6", "h10", "h6s"):
        T = g[f"{tag}_action"].shape[0]
        bits = g[f"{tag}_bits"].astype(np.uint32)
        full = bits | (g[f"{tag}_done_HA"].astype(np.uint32) << 16) | ((bits != 0).astype(np.uint32) << 18)
        prev = np.concatenate([[0], full[:-1]]).astype(np.uint32)
        pose = np.ascontiguousarray(g[f"{tag}_pre_rpose"][:, :3])
        pid = np.ascontiguousarray(g[f"{tag}_pre_path_idx"].astype(np.int32))
        lts = np.ascontiguousarray(g[f"{tag}_pre_loc_to_start"].astype(np.uint8))
        scripted = np.zeros(T, dtype=np.uint8); action = np.zeros((T, 2)); stuck = ctypes.c_uint8(0)
        assert lib.host_soft_reset_replay(T, p(prev), p(pose), p(pid), p(lts), p(tab), tab.shape[2], p(lens), p(scripted), p(action),
                                          ctypes.byref(stuck)) == 0
        soft = g[f"{tag}_soft"].astype(bool)
        assert np.array_equal(scripted.astype(bool), soft), np.argwhere(scripted.astype(bool) != soft)[:5]
        assert np.array_equal(action[soft], g[f"{tag}_action"][soft])
        assert stuck.value == 0
        a = g[f"{tag}_action"][soft]
        seen |= {"backup" if r[0] == -0.3 else "creep" if r[0] == 0.5 else "turn" if r[1] != 0 else "wait" for r in a}
    assert seen == {"wait", "backup", "turn", "creep"}


def test_shard_envs_partitions_exactly():
    from dr_mpc_b200.parallel import shard_envs
    for total, world in ((16384, 8), (65536, 8), (1000, 3), (5, 8), (16384, 1)):
        spans = [shard_envs(total, world, r) for r in range(world)]
        assert sum(c for _, c in spans) == total
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and spans[0][0] == 0
    with pytest.raises(ValueError):
        shard_envs(10, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as O
import dr_mpc_b200 as d
from dr_mpc_b200.parallel import init_distributed, shard_envs, episode_stats_from_step, reduce_episode_stats
rank, world, _ = init_distributed("gloo")
TOTAL, NH, STEPS = 48, 8, 6
tab, lens = d.paths.pack_paths(d.paths.continuous_task_paths())
off, cnt = shard_envs(TOTAL, world, rank)
env = O.EnvRef(cnt, NH, tab, lens, env_id_offset=off)     # the CPU oracle stands in for the per-GPU engine shard
a = env.reset(nthreads=1)
C = torch.zeros(16, dtype=torch.int64); S = torch.zeros(4, dtype=torch.float64)
for t in range(STEPS):
    a = env.step(a["mpc_actions"][:, :2].copy(), nthreads=1)
    c, s = episode_stats_from_step(torch.from_numpy(a["done_bits"].astype(np.int64)), torch.from_numpy(a["rewards"]))
    C += c; S += s
reduce_episode_stats(C, S)
np.savez(os.path.join(sys.argv[2], f"shard{rank}.npz"), hpos=a["hpos"], hgoal=a["hgoal"], rewards=a["rewards"], C=C.numpy(), S=S.numpy(), off=off, cnt=cnt)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_equals_single_process(oracle, tmp_path):
    """N>1 path on CPU (gloo, world_size 2): each rank steps its own contiguous env shard (no step-path collective),
    statistics are all-reduced; the union must equal the single-process run bit for bit (RNG keyed by global env id)."""
    import torch
    import dr_mpc_b200 as d
    from dr_mpc_b200.parallel import episode_stats_from_step
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    import socket
    with socket.socket() as sk:                  # a free rendezvous port instead of a fixed one
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), REPO, str(tmp_path)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    TOTAL, NH, STEPS = 48, 8, 6
    tab, lens = d.paths.pack_paths(d.paths.continuous_task_paths())
    ref = oracle.EnvRef(TOTAL, NH, tab, lens)
    a = ref.reset(nthreads=2)
    C = torch.zeros(16, dtype=torch.int64); S = torch.zeros(4, dtype=torch.float64)
    for t in range(STEPS):
        a = ref.step(a["mpc_actions"][:, :2].copy(), nthreads=2)
        c, s = episode_stats_from_step(torch.from_numpy(a["done_bits"].astype(np.int64)), torch.from_numpy(a["rewards"]))
        C += c; S += s
    for r in range(2):
        z = np.load(tmp_path / f"shard{r}.npz")
        sl = slice(int(z["off"]), int(z["off"]) + int(z["cnt"]))
        assert np.array_equal(z["hpos"], a["hpos"][sl]) and np.array_equal(z["hgoal"], a["hgoal"][sl]) and np.array_equal(z["rewards"], a["rewards"][sl])
        assert np.array_equal(z["C"], C.numpy()) and np.allclose(z["S"], S.numpy(), rtol=1e-12)
    assert C[0] == TOTAL * STEPS


# Golden solves on which the host build of the product header takes another LM branch than the reference did (explicit list;
# 1 of 583). F10 #8 is a FAILING first solve after a reset whose accept / reject sequence sits on a threshold: the host build
# (g++, no FMA contraction, another operation order than the oracle) finds an acceptable step where the reference raised and
# retried. The oracle and the cooperative CUDA kernel reproduce the reference on it; the test below shows the reference's own
# sequence flips under a 1e-13 perturbation of the input.
HOST_ALLOWED_DIVERGENT = {"F10": (8,)}


@pytest.mark.parametrize("K", [4, 10, 20, 40])
def test_device_mpc_header_on_host_matches_reference_golden(golden_mpc, K):
    """PRODUCT code on the CPU: the per-problem LM routine of dr-mpc_b200/csrc/mpc_device.cuh (the semantic reference the
    cooperative kernel is tested against on the GPU, and what DRE_MPC_SERIAL=1 runs there), compiled for the host, on every
    golden solve the reference's own MPC.act produced -- teacher-forced warm starts, all four paths, theta-wrap node, end-of-path
    clamping -- and, for K = 4 / 10, on the forced-failure solves that went through `pose_error_scaling *= 5` (mpc.py:188-190).
    Same bar as the GPU test: every solve within 1e-6, outer-iteration and retry counts equal, except the explicitly listed
    chaotic solve (HOST_ALLOWED_DIVERGENT)."""
    import dr_mpc_b200._lib as L
    from dr_mpc_b200.mpc import MPC_STATUS_DTYPE
    hh = os.path.join(REPO, "tests", "host_harness")
    so = os.path.join(hh, "libmpc_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
                           "-o", so, os.path.join(hh, "mpc_host.cpp")])
    lib = ctypes.CDLL(so)
    g = golden_mpc
    _, tab, lens = mpc_path_table(g)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    prm = L.MpcParams(K, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)
    for tag in ("K", "F") if K in (4, 10) else ("K",):
        k = f"{tag}{K}"
        n = g[f"{k}_u"].shape[0]
        wT = np.ascontiguousarray(g[f"{k}_warm_T"], dtype=np.float64).copy()
        wu = np.ascontiguousarray(g[f"{k}_warm_u"], dtype=np.float64).copy()
        it0 = np.ascontiguousarray(g[f"{k}_it0"]).astype(np.uint8)
        pid = np.ascontiguousarray(g[f"{k}_path_id"], dtype=np.int32)
        idx = np.ascontiguousarray(g[f"{k}_interp"], dtype=np.float64)
        pose = np.ascontiguousarray(g[f"{k}_pose"], dtype=np.float64)
        u = np.zeros((n, 2 * K))
        st = np.zeros(n, dtype=MPC_STATUS_DTYPE)
        assert lib.host_mpc_act(ctypes.byref(prm), n, p(tab), tab.shape[2], p(lens), p(pid), p(idx), p(pose), p(wT), p(wu), p(it0), p(u), p(st)) == 0
        d = np.abs(u - g[f"{k}_u"]).max(axis=1)
        ok = np.ones(n, dtype=bool)
        ok[list(HOST_ALLOWED_DIVERGENT.get(k, ()))] = False
        print(f"{k}: {n} solves, max|du| = {d[ok].max():.2e}, listed divergent: {np.nonzero(~ok)[0].tolist()}")
        assert d[ok].max() <= 1e-6, (k, np.nonzero(d > 1e-6)[0], d.max())
        assert np.array_equal(st["iterations"][ok], g[f"{k}_iters"][ok]) and (st["term"] > 0).all()
        if tag == "F":
            assert np.array_equal(st["retries"][ok], g[f"{k}_retries"][ok]) and (st["retries"][ok] >= 1).all()
        assert np.allclose(st["final_cost"][ok], g[f"{k}_cost"][ok], rtol=1e-6, atol=1e-9)
        for i in np.nonzero(~ok)[0]:
            # a listed solve must be one on which the REFERENCE's own decision sequence is unstable: the oracle (which reproduces
            # the golden on it) flips its LM branch under 1e-13 perturbations of the pose
            from oracle import oracle as O
            rng, flips = np.random.default_rng(int(i)), 0
            for _ in range(60):
                mb = O.MpcBatch(1, O.mpc_params(K=K), tab, lens)
                mb.warm_T[:] = g[f"{k}_warm_T"][i:i + 1]; mb.warm_u[:] = g[f"{k}_warm_u"][i:i + 1]; mb.iteration0[:] = g[f"{k}_it0"][i:i + 1]
                _, so = mb.act(pid[i:i + 1], idx[i:i + 1], pose[i:i + 1] + rng.uniform(-1e-13, 1e-13, (1, 3)))
                flips += int(so["retries"][0] != g[f"{k}_retries"][i] or so["iterations"][0] != g[f"{k}_iters"][i])
            assert flips >= 6, (k, int(i), flips)             # measured: 70 of 200


def test_robot_half_blocks_on_host_match_reference_rollout(golden_env, golden_mpc):
    """PRODUCT code on the CPU: the robot / path building blocks pt_step_kernel is made of (dr-mpc_b200/csrc/env_device.cuh:
    unicycle_step, localize_on_path, pt_state_gen), compiled for the host, on all 640 transitions the reference's own
    HAAndPTEnv.step recorded: post pose = exact-arc step of (pre pose, executed action) <= 1e-12 (agent.py:169-176,
    unicycle.py:89-126); 100-d PT state at the post pose <= 1e-9 (path_tracking_core.py:21-81,94-145). Measured: 4e-16 / 9e-16."""
    hh = os.path.join(REPO, "tests", "host_harness")
    so = os.path.join(hh, "libenv_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
                           "-o", so, os.path.join(hh, "env_host.cpp")])
    lib = ctypes.CDLL(so)
    g = golden_env
    _, tab, lens = mpc_path_table(golden_mpc)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    total = 0
    for tag in ("h6", "h10", "h20"):
        pre = np.ascontiguousarray(g[f"{tag}_pre_rpose"][:, :3])
        post = np.ascontiguousarray(g[f"{tag}_post_rpose"][:, :3])
        act = np.ascontiguousarray(g[f"{tag}_out_action"])
        T = pre.shape[0]
        total += T
        out = np.zeros((T, 3))
        assert lib.host_unicycle(T, p(pre), p(act), ctypes.c_double(0.25), p(out)) == 0
        assert np.abs(out - post).max() <= 1e-12, (tag, np.abs(out - post).max())
        pid = np.ascontiguousarray(g[f"{tag}_post_path_idx"].astype(np.int32))
        lts = np.ascontiguousarray(g[f"{tag}_post_loc_to_start"].astype(np.uint8))
        idx, pt = np.zeros(T), np.zeros((T, 100))
        assert lib.host_pt_state(T, p(post), p(pid), p(lts), p(tab), tab.shape[2], p(lens), 20, 5, p(idx), p(pt)) == 0
        assert np.abs(pt - g[f"{tag}_out_pt_state"]).max() <= 1e-9, (tag, np.abs(pt - g[f"{tag}_out_pt_state"]).max())
        assert (idx >= 0).all() and (idx <= lens[pid] - 1).all()                    # a fractional node index of the env's own path
    assert total == 640
