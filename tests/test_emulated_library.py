"""CPU: the WHOLE product library on an emulated GPU. tests/host_harness/gen_emu.py compiles every translation unit of
dr-mpc_b200/csrc -- the C ABI, the host pipelines and every CUDA kernel -- for the host after a handful of mechanical rewrites
(launch syntax, dynamic shared memory, two PTX lines), against a SIMT emulator (simt_emu.h: a CTA's threads are coroutines, warp /
CTA collectives are barrier exchanges) and a stand-in for the CUDA runtime calls of the host code (fake_cudart.h). The tests below
call the REAL `dre_*` entry points of that build and hold them to the GPU suite's bars on the reference's golden fixtures: what the
B200 executes -- pt_step_kernel, ha_step_kernel with both fused ORCA passes, mpc_coop_kernel, env_reset_kernel, cvmm_filter_kernel,
the soft-reset machine, the chunk-flag host pipeline -- executes here, thread for thread.

This is test infrastructure (nothing of it is shipped, the product path is the CUDA build and fails loudly without it); its point
is that the `-m "not gpu"` run already exercises the product's own kernels end to end. The B200 runs of the same fixtures are in
tests/test_gpu_*.py."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ORCA_KEYS, REPO, mpc_path_table
import emu_driver as E_


@pytest.fixture(scope="module")
def L():
    return E_.lib()


def test_emulated_build_is_the_whole_c_abi(L):
    """every function include/dre.h declares is exported by the emulated build too (it IS the product's source, all six units)"""
    hdr = open(os.path.join(REPO, "include", "dre.h")).read()
    names = sorted(set(re.findall(r"\b(dre_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), n
    assert L.dre_version() >= 100


@pytest.mark.parametrize("lanes", [0, 8, 32])
def test_orca_predict_kernels_bit_exact_on_golden(L, golden_orca, lanes):
    """dre_orca_predict -> orca_predict_kernel<G> (auto / 8 / 32 lanes per agent) on all 248 golden crowds"""
    from dr_mpc_b200.config import EngineConfig
    g = golden_orca
    for key in ORCA_KEYS:
        hpos, hvel = np.ascontiguousarray(g[key + "_hpos"]), np.ascontiguousarray(g[key + "_hvel"])
        goal, rpos = np.ascontiguousarray(g[key + "_goal"]), np.ascontiguousarray(g[key + "_rpos"])
        n, Nh, _ = hpos.shape
        prm = EngineConfig(group_lanes=lanes).orca_params()
        v = np.zeros((n, Nh, 2), np.float32)
        E_.check(L.dre_orca_predict(ctypes.byref(prm), n, Nh, E_.ptr(hpos), E_.ptr(hvel), E_.ptr(goal), E_.ptr(rpos), None, E_.ptr(v), None, None))
        assert np.array_equal(v.view(np.uint32), g[key + "_vnew"].view(np.uint32)), (key, lanes)


@pytest.mark.parametrize("key", ["K4", "F4", "K10", "F10", "K20", "K40"])
def test_mpc_act_host_on_golden(L, golden_mpc, key):
    """dre_mpc_create / set_paths / set_warm / act_host -> the persistent mpc_coop_kernel on a multi-CTA grid"""
    import dr_mpc_b200._lib as LL
    from dr_mpc_b200.mpc import MPC_STATUS_DTYPE
    g = golden_mpc
    K = int(key[1:])
    _, tab, lens = mpc_path_table(g)
    n = g[f"{key}_u"].shape[0]
    prm = LL.MpcParams(K, 0.25, 1.0, 1.0, 0.2, 1500, 1e-4, 1e-4, 0, 20)
    h = ctypes.c_void_p()
    E_.check(L.dre_mpc_create(ctypes.byref(prm), n, ctypes.byref(h)))
    try:
        E_.check(L.dre_mpc_set_paths(h, E_.ptr(tab), 4, tab.shape[2], E_.ptr(lens)))
        wT, wu = np.ascontiguousarray(g[f"{key}_warm_T"], dtype=np.float64), np.ascontiguousarray(g[f"{key}_warm_u"], dtype=np.float64)
        it0 = np.ascontiguousarray(g[f"{key}_it0"]).astype(np.uint8)
        E_.check(L.dre_mpc_set_warm(h, E_.ptr(wT), E_.ptr(wu), E_.ptr(it0)))
        pid, idx = np.ascontiguousarray(g[f"{key}_path_id"], dtype=np.int32), np.ascontiguousarray(g[f"{key}_interp"], dtype=np.float64)
        pose = np.ascontiguousarray(g[f"{key}_pose"], dtype=np.float64)
        u, st = np.zeros((n, 2 * K)), np.zeros(n, dtype=MPC_STATUS_DTYPE)
        E_.check(L.dre_mpc_act_host(h, E_.ptr(pid), E_.ptr(idx), E_.ptr(pose), E_.ptr(u), E_.ptr(st)))
    finally:
        L.dre_mpc_destroy(h)
    ok = np.ones(n, dtype=bool)
    if key == "F10":
        ok[8] = False                          # the listed chaotic solve (tests/test_host.py: HOST_ALLOWED_DIVERGENT)
    d = np.abs(u - g[f"{key}_u"]).max(1)
    assert d[ok].max() <= 1e-6 and np.array_equal(st["iterations"][ok], g[f"{key}_iters"][ok]) and (st["term"] > 0).all()
    if key[0] == "F":
        assert np.array_equal(st["retries"][ok], g[f"{key}_retries"][ok])


@pytest.mark.parametrize("tag,Nh,count", [("h6", 6, 500), ("h10", 10, 90), ("h20", 20, 50)])
def test_env_step_matches_reference_teacher_forced(golden_env, tag, Nh, count):
    """tests/test_gpu_env.py::test_step_matches_reference_teacher_forced on the emulated build: every recorded transition of the
    reference's own HAAndPTEnv.step is one env of a batch -- upload the reference's pre-step state, ONE dre_env_step (pt_step,
    ha_step with both ORCA passes, the MPC solve next to it), compare observations / rewards / done reasons / post-step state."""
    g = golden_env
    T = min(count, g[f"{tag}_out_action"].shape[0])
    sel = slice(0, T)
    env = E_.EmuEnv(T, num_humans=Nh, auto_path_switch=False)
    env.reset()
    env.upload(g, tag, "pre", sel)
    o = env.step(g[f"{tag}_out_action"][sel])
    st = env.state
    assert np.array_equal(st["hvel"].view(np.uint32), g[f"{tag}_post_hvel"][sel].view(np.uint32))
    assert np.abs(st["hpos"] - g[f"{tag}_post_hpos"][sel]).max() <= 1e-12
    assert np.abs(st["hhist"] - g[f"{tag}_post_hhist"][sel]).max() <= 1e-12
    assert np.array_equal(st["hvalid"], g[f"{tag}_post_hvalid"][sel])
    assert np.abs(st["rpose"][:, :3] - g[f"{tag}_post_rpose"][sel][:, :3]).max() <= 1e-12
    assert np.abs(st["rhist"][:, :, :3] - g[f"{tag}_post_rhist"][sel][:, :, :3]).max() <= 1e-12
    assert np.abs(st["rpastv"] - g[f"{tag}_post_rpastv"][sel]).max() <= 1e-15
    assert np.abs(st["gtime"] - g[f"{tag}_post_gtime"][sel]).max() <= 1e-12
    assert np.array_equal(st["loc_to_start"], g[f"{tag}_post_loc_to_start"][sel])
    changed = np.abs(g[f"{tag}_post_hgoal"][sel] - g[f"{tag}_pre_hgoal"][sel]).max(-1) > 0
    assert np.array_equal(st["hgoal"][~changed], g[f"{tag}_post_hgoal"][sel][~changed])
    for name, tol in (("pt_state", 1e-9), ("robot_node", 1e-9), ("past_robot_vel", 1e-15), ("percent_episode", 1e-12),
                      ("spatial_edges", 1e-9), ("human_valid", 0), ("detected_humans", 0)):
        ref = g[f"{tag}_out_{name}"][sel].reshape(o[name].shape)
        assert np.abs(o[name] - ref).max() <= tol, name
    assert np.abs(o["rewards"] - g[f"{tag}_out_rewards"][sel]).max() <= 1e-9
    for col, name in enumerate(("disturbance_v", "disturbance_th", "path_advancement", "deviation")):
        assert np.abs(o["info"][:, col] - g[f"{tag}_out_{name}"][sel]).max() <= 1e-9, name
    bits = g[f"{tag}_out_bits"][sel].astype(np.int64)
    assert np.array_equal(o["done_bits"] & 0xFF, bits & 0xFF)
    fl = env.flags()
    assert np.array_equal(fl["done_HA"], g[f"{tag}_out_done_HA"][sel].astype(bool)) and np.array_equal(fl["done_PT"], g[f"{tag}_out_done_PT"][sel].astype(bool))
    assert np.array_equal(fl["done"], (bits & 0xFF) != 0) and np.array_equal(fl["success"], (bits & 16) != 0)
    d = np.abs(o["mpc_actions"] - g[f"{tag}_out_mpc_actions"][sel]).max(1)
    assert d.max() <= 1e-6, d.max()


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10), ("h20", 20)])
def test_env_reset_matches_reference(golden_env, tag, Nh):
    """dre_env_reset: env_reset_kernel, the first MPC solve and the six warm-start passes of ha_step_kernel, from the reference's spawn"""
    g = golden_env
    env = E_.EmuEnv(3, num_humans=Nh, end_goal_changing=False)
    o = env.reset(np.broadcast_to(g[f"{tag}_reset_hpos"], (3, Nh, 2)).copy(), np.broadcast_to(g[f"{tag}_reset_hgoal"], (3, Nh, 2)).copy())
    st = env.state
    assert np.array_equal(st["hvel"][0].view(np.uint32), g[f"{tag}_postreset_hvel"].view(np.uint32))
    assert np.abs(st["hpos"][0] - g[f"{tag}_postreset_hpos"]).max() <= 1e-12
    assert np.abs(st["hhist"][0] - g[f"{tag}_postreset_hhist"]).max() <= 1e-12
    assert np.abs(st["rhist"][0, :, :3] - g[f"{tag}_postreset_rhist"][:, :3]).max() <= 1e-12
    assert st["rvalid"][0] == g[f"{tag}_postreset_rvalid"] and np.array_equal(st["hvalid"][0], g[f"{tag}_postreset_hvalid"])
    for name, tol in (("pt_state", 1e-9), ("robot_node", 1e-9), ("past_robot_vel", 0), ("percent_episode", 0), ("spatial_edges", 1e-9),
                      ("human_valid", 0), ("mpc_actions", 1e-6)):
        ref = g[f"{tag}_reset_{name}"].reshape(o[name][0].shape)
        assert np.abs(o[name][0] - ref).max() <= tol, name
    assert (o["pt_state"][0] == o["pt_state"][2]).all()


@pytest.mark.parametrize("tag,Nh,count", [("h6", 6, 10 ** 6), ("h10", 10, 10 ** 6), ("h6s", 6, 10 ** 6)])
def test_soft_reset_machine_follows_reference(tag, Nh, count):
    """tests/test_gpu_soft_reset.py on the emulated build: the device-side machine (pt_step_kernel: soft_reset_decide) has to walk,
    from its own done bits, through the waits / back-ups / turns / creeps the reference's HAAndPTEnv.soft_reset issued; the physical
    state is teacher-forced before every call, the machine's state is not."""
    g = np.load(os.path.join(GOLDEN, "soft_reset.npz"))
    T = min(count, g[f"{tag}_action"].shape[0])
    env = E_.EmuEnv(1, num_humans=Nh, soft_reset_on_device=True)
    env.reset()
    kinds = {"caller": 0, "wait": 0, "backup": 0, "turn": 0, "creep": 0}
    for t in range(T):
        env.upload(g, tag, "pre", slice(t, t + 1))
        soft, ref_a = bool(g[f"{tag}_soft"][t]), g[f"{tag}_action"][t]
        o = env.step(np.array([[0.777, -0.555]]) if soft else ref_a[None])      # a scripted call gets a poison action from the caller
        fl = env.flags()
        assert bool(fl["soft_reset"][0]) == soft, (t, "scripted flag")
        assert np.array_equal(o["actions_used"][0], ref_a), (t, o["actions_used"][0], ref_a)
        assert not fl["soft_reset_stuck"][0]
        assert (int(o["done_bits"][0]) & 0xFF) == (int(g[f"{tag}_bits"][t]) & 0xFF), t
        assert np.abs(o["rewards"][0] - g[f"{tag}_rewards"][t]).max() <= 1e-9, t
        assert np.abs(env.state["rpose"][0, :3] - g[f"{tag}_post_rpose"][t, :3]).max() <= 1e-12
        assert np.array_equal(env.state["hvel"][0].view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32))
        kinds["caller" if not soft else "backup" if ref_a[0] == -0.3 else "creep" if ref_a[0] == 0.5 else "turn" if ref_a[1] != 0 else "wait"] += 1
    print(tag, T, "calls", kinds)
    assert kinds["caller"] > 0 and kinds["wait"] > 0
    if tag == "h6":
        assert kinds["turn"] > 0 and kinds["creep"] > 0
    if tag == "h6s":
        assert kinds["backup"] > 20


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_cvmm_filter_matches_reference(tag, Nh):
    """dre_env_cvmm_filter -> cvmm_filter_kernel (one warp per env, lane = candidate) on the recorded queries of the reference's
    cvmm_diverse_safety_pipeline: per-candidate safety flags and the selected action exact"""
    g = np.load(os.path.join(GOLDEN, "cvmm.npz"))
    Q = g[f"{tag}_action"].shape[0]
    sel = slice(0, Q)
    env = E_.EmuEnv(Q, num_humans=Nh)
    env.reset()
    for name in E_.STATE:
        env.state[name][...] = np.asarray(g[f"{tag}_pre_{name}"][sel]).reshape(env.state[name].shape)
    before = {k: v.copy() for k, v in env.state.items()}
    out, info = env.cvmm(g[f"{tag}_action"][sel], int(g["lookahead_steps"]), g["action_set"])
    assert np.array_equal(info["candidate_safe"], g[f"{tag}_cand_safe"][sel])
    rnd = g[f"{tag}_random"][sel].astype(bool)
    assert np.array_equal(info["random"].astype(bool), rnd)
    assert np.array_equal(out[~rnd], g[f"{tag}_safe_action"][sel][~rnd])
    kept = info["chosen"] == -1
    assert np.array_equal(kept, g[f"{tag}_cand_safe"][sel][:, 0].astype(bool)) and np.array_equal(out[kept], g[f"{tag}_action"][sel][kept])
    for k, v in env.state.items():
        assert np.array_equal(v, before[k]), k                                   # a pure query


@pytest.mark.parametrize("chunks", [1, 3])
def test_step_host_pipeline_equals_device_step(chunks):
    """dre_env_step_host (one crowd-half launch whose CTAs raise chunk flags in mapped host memory, the host thread polling them
    and queueing each chunk's copy) returns the device step's slab bit for bit, for a ragged chunking"""
    Ee, Nh = 75, 7
    a_env, b_env = E_.EmuEnv(Ee, num_humans=Nh), E_.EmuEnv(Ee, num_humans=Nh)
    b_env.set_host_chunks(chunks)
    act = a_env.reset()["mpc_actions"][:, :2].copy()
    b_env.reset()
    for t in range(3):
        oa = a_env.step(act)
        ob = b_env.step_host(act)
        for name in oa:
            assert np.array_equal(oa[name].view(np.uint8), ob[name].view(np.uint8)), (t, name)
        for name in ("hpos", "hvel", "hgoal", "rpose", "hhist"):
            assert np.array_equal(a_env.state[name], b_env.state[name]), (t, name)
        act = oa["mpc_actions"][:, :2].copy()


def test_free_run_double_buffer_masked_reset_and_sharding():
    """a short free run on the emulated build: outputs finite, the slab alternates (S of step t-1 intact while S' of step t exists),
    a masked reset touches only its envs, and two shards keyed by the global env id equal the unsharded batch bit for bit"""
    Ee, Nh = 48, 6
    env = E_.EmuEnv(Ee, num_humans=Nh)
    S0 = env.reset()
    first = {k: v.copy() for k, v in S0.items()}
    a = S0["mpc_actions"][:, :2].copy()
    o1 = env.step(a)
    assert o1 is not S0 and all(np.array_equal(S0[k], first[k]) for k in first)      # the reset slab survived the step
    for _ in range(4):
        o1 = env.step(o1["mpc_actions"][:, :2].copy())
    assert np.isfinite(o1["rewards"]).all() and np.isfinite(o1["spatial_edges"]).all() and (env.state["step_count"] == 5).all()
    keep = {k: v.copy() for k, v in env.state.items()}
    mask = np.zeros(Ee, np.uint8)
    mask[::3] = 1
    env.reset(env_mask=mask)
    on = mask.astype(bool)
    assert (env.state["step_count"][on] == 0).all() and np.array_equal(env.state["step_count"][~on], keep["step_count"][~on])
    assert np.array_equal(env.state["hpos"][~on], keep["hpos"][~on]) and not np.array_equal(env.state["hpos"][on], keep["hpos"][on])
    # sharding: envs [0, 20) and [20, 48) as separate handles with env_id_offset == one handle of 48
    whole = E_.EmuEnv(Ee, num_humans=Nh, seed=7)
    parts = [E_.EmuEnv(20, num_humans=Nh, seed=7, env_id_offset=0), E_.EmuEnv(28, num_humans=Nh, seed=7, env_id_offset=20)]
    ow = whole.reset()
    op = [p.reset() for p in parts]
    for _ in range(3):
        ow = whole.step(ow["mpc_actions"][:, :2].copy())
        op = [p.step(o["mpc_actions"][:, :2].copy()) for p, o in zip(parts, op)]
    for name in ("spatial_edges", "rewards", "pt_state", "done_bits"):
        assert np.array_equal(ow[name], np.concatenate([o[name] for o in op])), name


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_sensor_restriction_matches_reference(tag, Nh):
    """tests/test_gpu_sensor.py on the emulated build: humans beyond sensor_range lose their history and the observation lists the
    visible ones first (human_avoidance_env.py:53-72,102-105,161-164), step and reset"""
    g = np.load(os.path.join(GOLDEN, "sensor_restriction.npz"))
    T = g[f"{tag}_out_action"].shape[0]
    env = E_.EmuEnv(T, num_humans=Nh, robot_sensor_restriction=True, auto_path_switch=False)
    env.reset()
    env.upload(g, tag, "pre")
    o = env.step(g[f"{tag}_out_action"])
    st = env.state
    assert np.array_equal(st["hvalid"], g[f"{tag}_post_hvalid"]) and np.abs(st["hhist"] - g[f"{tag}_post_hhist"]).max() <= 1e-12
    assert np.array_equal(o["detected_humans"], g[f"{tag}_out_detected_humans"].reshape(-1))
    assert np.array_equal(o["human_valid"], g[f"{tag}_out_human_valid"].reshape(o["human_valid"].shape))
    assert np.abs(o["spatial_edges"] - g[f"{tag}_out_spatial_edges"].reshape(o["spatial_edges"].shape)).max() <= 1e-9
    assert np.abs(o["rewards"] - g[f"{tag}_out_rewards"]).max() <= 1e-9
    assert (st["hvalid"].sum(1) == 0).any() and ((g[f"{tag}_pre_hvalid"] > 0) & (st["hvalid"] == 0)).any()
    env2 = E_.EmuEnv(3, num_humans=Nh, robot_sensor_restriction=True)
    o2 = env2.reset(np.broadcast_to(g[f"{tag}_reset_hpos"], (3, Nh, 2)).copy(), np.broadcast_to(g[f"{tag}_reset_hgoal"], (3, Nh, 2)).copy())
    assert np.array_equal(o2["detected_humans"], np.full(3, g[f"{tag}_reset_detected_humans"]))
    assert np.array_equal(o2["human_valid"][0], g[f"{tag}_reset_human_valid"])
    assert np.abs(o2["spatial_edges"][0] - g[f"{tag}_reset_spatial_edges"]).max() <= 1e-9


@pytest.mark.parametrize("Ee,Nh,variant", [(64, 20, {}), (24, 50, {}), (40, 12, {"robot_visible": False}), (40, 16, {"neighbor_dist": 3.0}),
                                           (32, 6, {"horizon": 10}), (10, 6, {"horizon": 40}),
                                           (64, 6, {"soft_reset_on_device": True, "static_humans": ((-4.0, 0.0), (0.0, 0.0))}),
                                           (40, 10, {"robot_sensor_restriction": True})])
def test_free_running_step_matches_oracle(oracle, golden_mpc, Ee, Nh, variant):
    """tests/test_gpu_env.py::test_step_matches_oracle_at_scale on the emulated build: states the free-running engine itself
    reaches (Philox spawning and goal resampling, path switches, the soft-reset machine, static humans, the sensor restriction,
    the 3 m neighbour cut-off variant, horizons on 16-lane and two-warp groups) are copied into oracle/env_ref.c, both take the
    same step: velocities bit-exact, positions / goals 1e-12, rewards / observations 1e-9, done reasons and executed actions equal."""
    _, tab, lens = mpc_path_table(golden_mpc)
    env = E_.EmuEnv(Ee, num_humans=Nh, **variant)
    o = env.reset()
    ov = dict(variant)
    ref_kw = {k: ov.pop(k) for k in ("neighbor_dist", "static_humans") if k in ov}
    if "horizon" in ov: ref_kw["K"] = ov.pop("horizon")
    if "soft_reset_on_device" in ov: ref_kw["soft_reset"] = ov.pop("soft_reset_on_device")
    if "robot_sensor_restriction" in ov: ref_kw["sensor_restriction"] = ov.pop("robot_sensor_restriction")
    ref = oracle.EnvRef(Ee, Nh, tab, lens, **ref_kw, **{k: int(v) for k, v in ov.items()})
    ref.reset()
    rng = np.random.default_rng(7)
    checked = lm_div = scripted = 0
    steps = 90 if variant.get("soft_reset_on_device") else 26
    for t in range(steps):
        a = o["mpc_actions"][:, :2] + rng.normal(size=(Ee, 2)) * 0.15
        a[:, 0] = a[:, 0].clip(0, 1); a[:, 1] = a[:, 1].clip(-1, 1)
        probe = t % 5 == 3
        if probe:
            for k in ("hpos", "hgoal", "hhist", "hvel", "hstatic", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
                      "path_idx", "loc_to_start", "step_count", "sr_state"):
                ref.a[k][...] = env.state[k].reshape(ref.a[k].shape)
            ref.a["done_bits"][...] = o["done_bits"]
            ref.a["reset_epoch"][...] = 1
            ref.a["warm_T"][...], ref.a["warm_u"][...], ref.a["iteration0"][...] = env.mpc_get_warm()
        o = env.step(a)
        if not probe:
            continue
        r = ref.step(a, nthreads=0)
        st = env.state
        assert np.array_equal(st["hvel"].view(np.uint32), r["hvel"].view(np.uint32)), t
        assert np.abs(st["hpos"] - r["hpos"]).max() <= 1e-12 and np.abs(st["hgoal"] - r["hgoal"]).max() <= 1e-12
        assert np.abs(st["rpose"][:, :3] - r["rpose"][:, :3]).max() <= 1e-12 and np.array_equal(st["path_idx"], r["path_idx"])
        assert np.array_equal(o["done_bits"], r["done_bits"]) and np.array_equal(o["actions_used"], r["actions_used"]), t
        if variant.get("soft_reset_on_device"):
            assert np.array_equal(o["done_flags"][:, 12], r["soft_flags"][:, 0]) and np.array_equal(st["sr_state"], r["sr_state"]), t
            scripted += int(r["soft_flags"][:, 0].sum())
        assert np.array_equal(st["hvalid"], r["hvalid"]) and np.array_equal(o["detected_humans"], r["detected_humans"]), t
        assert np.abs(o["rewards"] - r["rewards"]).max() <= 1e-9
        for name in ("pt_state", "robot_node", "past_robot_vel", "spatial_edges", "human_valid"):
            assert np.abs(o[name] - r[name].reshape(o[name].shape)).max() <= 1e-9, (t, name)
        stat = o["mpc_status"]
        same = (stat[:, 0] == r["mpc_status"]["iterations"]) & (stat[:, 2] == r["mpc_status"]["lm_solves"]) & (stat[:, 3] == r["mpc_status"]["cost_evals"])
        assert np.abs(o["mpc_actions"] - r["mpc_actions"]).max(1)[same].max() <= 1e-6
        checked += Ee
        lm_div += int((~same).sum())
    assert checked >= 5 * Ee and lm_div <= max(2, 0.02 * checked)
    if variant.get("soft_reset_on_device"):
        assert scripted > 10


@pytest.mark.parametrize("flag", ["orca_dedup", "orca_lookahead_prune"])
def test_orca_variants_are_bit_identical(flag):
    """the two opt-in ORCA variants against the reference's two full passes, free-running with goal resamplings: every output
    and state array equal bit for bit (tests/test_gpu_env.py: test_orca_dedup_is_bit_identical / test_lookahead_prune_is_bit_identical)"""
    Ee, Nh = 40, 10
    a_env, b_env = E_.EmuEnv(Ee, num_humans=Nh), E_.EmuEnv(Ee, num_humans=Nh, **{flag: True})
    oa = a_env.reset()
    b_env.reset()
    resampled = 0
    for t in range(40):
        act = oa["mpc_actions"][:, :2].copy()
        g0 = a_env.state["hgoal"].copy()
        oa, ob = a_env.step(act), b_env.step(act)
        resampled += int((a_env.state["hgoal"] != g0).any(-1).sum())
        for k in oa:
            assert np.array_equal(oa[k].view(np.uint8), ob[k].view(np.uint8)), (t, k)
        for k in ("hpos", "hvel", "hgoal", "hhist", "rpose", "path_idx"):
            assert np.array_equal(a_env.state[k], b_env.state[k]), (t, k)
    assert resampled > 2


@pytest.mark.parametrize("Ee,Nh,cap,dtype", [(64, 6, 150, np.float64), (40, 7, 1000, np.float64), (50, 5, 30, np.float32), (1100, 3, 5000, np.float64)])
def test_replay_ring_against_a_numpy_ring(Ee, Nh, cap, dtype):
    """dre_replay_insert / sample (replay_select_kernel + replay_gather_kernel) against the reference buffer's semantics written
    out in numpy (storage.py:54-133: rows in insertion order, the oldest rows fall out when full): masked inserts with the
    soft-reset machine on (scripted steps stay out), wrap-arounds, odd-width fields, an insert larger than the ring, float32 storage,
    1100 envs = two selection CTAs (the last-CTA scan)"""
    env = E_.EmuEnv(Ee, num_humans=Nh, soft_reset_on_device=True)
    rb = E_.EmuReplay(env, cap, dtype)
    S = {k: v.copy() for k, v in env.reset().items()}
    rows = []                                    # chronological list of (S, A, S', R, done) rows, the newest `cap` survive
    rng = np.random.default_rng(Ee)
    for t in range(12 if Ee < 1000 else 6):
        a = S["mpc_actions"][:, :2].copy()
        a[:, 1] = (a[:, 1] + 0.4 * np.sin(np.arange(Ee) * 0.37 + t)).clip(-1, 1)      # a sloppy policy: terminations happen
        o = env.step(a)
        mask = rng.uniform(size=Ee) < 0.8
        rb.insert(a, mask)
        keep = mask & (o["done_flags"][:, 12] == 0)
        for e in np.nonzero(keep)[0]:
            rows.append(({k: S[k][e].ravel().copy() for k in E_.OBS_KEYS}, a[e].copy(), {k: o[k][e].ravel().copy() for k in E_.OBS_KEYS},
                         o["rewards"][e, 2], float(o["done_flags"][e, 3] != 0)))
        assert rb.count()[2] == min(int(keep.sum()), cap)
        S = {k: v.copy() for k, v in o.items()}
    rows = rows[-cap:]
    valid, head, _ = rb.count()
    assert valid == len(rows)
    chron = rb.sample(np.arange(valid))
    for j, (s, a, sp, r, dn) in enumerate(rows):
        for k in E_.OBS_KEYS:
            assert np.array_equal(chron["S"][k][j], s[k].astype(dtype)), (j, k)
            assert np.array_equal(chron["S_prime"][k][j], sp[k].astype(dtype)), (j, k)
        assert np.array_equal(chron["actions"][j], a.astype(dtype)) and chron["rewards"][j, 0] == dtype(r) and chron["done"][j, 0] == dn
    # the ring's own storage holds the same rows at (head + j) % cap once full
    phys = (np.arange(valid) + (head if valid == cap else 0)) % cap
    assert np.array_equal(rb.ring["S_prime"]["spatial_edges"][phys], chron["S_prime"]["spatial_edges"])
    if dtype == np.float64:
        half = rb.sample(np.arange(valid), dtype=np.float32)
        assert np.array_equal(half["S"]["pt_state"], chron["S"]["pt_state"].astype(np.float32))


def test_switch_path_masked_observe_and_soft_reset_decide():
    """tests/test_gpu_api.py on the emulated build: the host-driven form of soft_reset's path switch (dre_env_switch_path +
    dre_env_observe_masked) equals a second engine put on the new path by hand; dre_env_soft_reset_decide(commit=0) predicts
    what the device machine does in the next step and leaves its state alone."""
    Ee = 40
    a_env = E_.EmuEnv(Ee, num_humans=6, auto_path_switch=False)
    o = a_env.reset()
    for _ in range(4):
        o = a_env.step(o["mpc_actions"][:, :2].copy())
    mask = np.zeros(Ee, bool)
    mask[1::2] = True
    pre = {k: v.copy() for k, v in o.items()}
    T0, u0, _ = a_env.mpc_get_warm()
    a_env.switch_path(mask)
    o2 = a_env.observe(solve_mpc=True, env_mask=mask)
    st = a_env.state
    assert (st["path_idx"][mask] == 1).all() and (st["path_idx"][~mask] == 0).all() and (st["loc_to_start"][mask] == 1).all()
    for k in pre:
        assert np.array_equal(o2[k][~mask], pre[k][~mask]), k
    T1, u1, it1 = a_env.mpc_get_warm()
    assert np.array_equal(T0[~mask], T1[~mask]) and (it1 == 0).all() and not np.array_equal(o2["pt_state"][mask], pre["pt_state"][mask])
    b_env = E_.EmuEnv(Ee, num_humans=6, auto_path_switch=False)
    b_env.reset()
    for k in a_env.state:
        b_env.state[k][...] = a_env.state[k]
    b_env.mpc_set_warm(T0, u0, np.ones(Ee, np.uint8))
    ob = b_env.observe(solve_mpc=True)
    assert np.array_equal(ob["pt_state"][mask], o2["pt_state"][mask]) and np.array_equal(ob["mpc_actions"][mask], o2["mpc_actions"][mask])
    env = E_.EmuEnv(96, num_humans=10, soft_reset_on_device=True)
    o = env.reset()
    seen = 0
    for t in range(70):
        sr0 = env.state["sr_state"].copy()
        scripted, act, stuck = env.soft_reset_decide(commit=False)
        assert np.array_equal(env.state["sr_state"], sr0)
        a = o["mpc_actions"][:, :2].copy()
        a[:, 1] = (a[:, 1] + 0.5 * np.sin(np.arange(96) * 0.37 + t)).clip(-1, 1)
        o = env.step(a)
        assert np.array_equal(env.flags()["soft_reset"], scripted), t
        assert np.array_equal(o["actions_used"][scripted], act[scripted]) and not stuck.any()
        seen += int(scripted.sum())
    assert seen > 5


def test_host_output_modes_cvmm_check_crowd_step_and_stats():
    """dre_env_set_host_outputs (float32 observations through cvt32_kernel = the float64 results rounded once; control fields
    only), dre_env_cvmm_check against the filter's verdict on the original action, dre_env_crowd_step (BASELINE config 4's
    ORCA-only step) and the episode statistics"""
    Ee, Nh = 45, 10
    envs = {m: E_.EmuEnv(Ee, num_humans=Nh) for m in ("full", "f32", "ctl")}
    envs["f32"].set_host_outputs(1)
    envs["ctl"].set_host_outputs(2)
    act = envs["full"].reset()["mpc_actions"][:, :2].copy()
    envs["f32"].reset(); envs["ctl"].reset()
    obs = ("pt_state", "robot_node", "past_robot_vel", "percent_episode", "spatial_edges", "human_valid", "detected_humans")
    for t in range(3):
        v = {m: e.step_host_mode(act) for m, e in envs.items()}
        for name in obs:
            assert v["f32"][name].dtype == np.float32 and name not in v["ctl"]
            assert np.array_equal(v["f32"][name], v["full"][name].astype(np.float32)), (t, name)
        for name in ("actions_used", "mpc_actions", "rewards", "info", "done_bits", "mpc_status", "done_flags"):
            for m in ("f32", "ctl"):
                assert np.array_equal(v[m][name].view(np.uint8), v["full"][name].view(np.uint8)), (t, m, name)
        for m in ("f32", "ctl"):
            assert np.array_equal(envs[m].out["spatial_edges"], envs["full"].out["spatial_edges"])      # the device slab stays float64
        act = v["full"]["mpc_actions"][:, :2].copy()
    env = envs["full"]
    safe, dist = env.cvmm_check(act, 8)
    _, info = env.cvmm(act, 8, np.array([[0.05, -0.8], [0.6, 0.0], [0.15, 0.2]]))
    assert np.array_equal(safe, info["original_safe"].astype(bool)) and np.isfinite(dist).all()
    c, s = env.stats()
    assert c[0] == 3 * Ee and np.isfinite(s).all()
    before = {k: env.state[k].copy() for k in ("hpos", "rpose", "step_count")}
    env.crowd_step(write_obs=True)
    assert not np.array_equal(env.state["hpos"], before["hpos"]) and np.array_equal(env.state["rpose"], before["rpose"])
    assert np.abs(env.state["hpos"] - (before["hpos"] + env.state["hvel"].astype(np.float64) * 0.25)).max() <= 1e-15


@pytest.mark.parametrize("Ee,Nh,K", [(1, 1, 4), (3, 64, 4), (37, 63, 1), (5, 5, 40), (2, 6, 33), (130, 3, 8)])
def test_edge_shapes_against_the_oracle(oracle, golden_mpc, Ee, Nh, K):
    """the corners of the supported range -- a single env with a single human, the largest crowd (64 humans: the shared-memory
    layout at its limit, the thread-per-agent ORCA fallback), horizons 1 / 33 / 40 (one lane, the two-warp groups), env counts that
    are no multiple of any tile -- free-running for a few steps, every step compared with oracle/env_ref.c from the same state"""
    _, tab, lens = mpc_path_table(golden_mpc)
    env = E_.EmuEnv(Ee, num_humans=Nh, horizon=K)
    o = env.reset()
    ref = oracle.EnvRef(Ee, Nh, tab, lens, K=K)
    ref.reset()
    for t in range(5):
        a = o["mpc_actions"][:, :2].copy()
        for k in ("hpos", "hgoal", "hhist", "hvel", "hstatic", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
                  "path_idx", "loc_to_start", "step_count", "sr_state"):
            ref.a[k][...] = env.state[k].reshape(ref.a[k].shape)
        ref.a["done_bits"][...] = o["done_bits"]
        ref.a["reset_epoch"][...] = 1
        ref.a["warm_T"][...], ref.a["warm_u"][...], ref.a["iteration0"][...] = env.mpc_get_warm()
        o = env.step(a)
        r = ref.step(a, nthreads=1)
        assert np.array_equal(env.state["hvel"].view(np.uint32), r["hvel"].view(np.uint32)), t
        assert np.abs(env.state["hpos"] - r["hpos"]).max() <= 1e-12 and np.array_equal(o["done_bits"], r["done_bits"])
        assert np.abs(o["rewards"] - r["rewards"]).max() <= 1e-9
        for name in ("pt_state", "robot_node", "spatial_edges", "human_valid"):
            assert np.abs(o[name] - r[name].reshape(o[name].shape)).max() <= 1e-9, (t, name)
        same = o["mpc_status"][:, 0] == r["mpc_status"]["iterations"]
        assert same.mean() >= 0.9 and np.abs(o["mpc_actions"] - r["mpc_actions"]).max(1)[same].max() <= 5e-6


@pytest.mark.parametrize("seed", [1, 5, 7, 23, 47, 52])
def test_fuzzed_configurations_against_the_oracle(seed):
    """tests/emu_fuzz.py: shape, horizon, neighbour range, robot visibility, static humans, sensor restriction, soft reset, goal
    policy, path switching, time limit and action noise drawn per seed; free-running, every third step compared with
    oracle/env_ref.c from the same state (400 seeds run clean when this was written: profiles/r2_emu_fuzz.txt)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_emu_fuzz", os.path.join(REPO, "tests", "emu_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    assert fz.one(seed, steps=25)
