"""CPU: the flat C env oracle (oracle/env_ref.c) against every transition recorded from the reference's own
HAAndPTEnv.step (unmodified Python on the shims): upload the pre-step state, step once, compare (teacher forcing)."""
import numpy as np
import pytest
from conftest import mpc_path_table

TAGS = [("h6", 6), ("h10", 10), ("h20", 20)]


def _load(env, g, tag, prefix, sel=slice(None)):
    for name in ("hpos", "hvel", "hgoal", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime", "path_idx", "loc_to_start", "warm_T", "warm_u"):
        a = g[f"{tag}_{prefix}_{name}"][sel]
        env.a[name][...] = a.reshape(env.a[name].shape)
    env.a["iteration0"][...] = g[f"{tag}_{prefix}_it0"][sel]
    env.a["hstatic"][...] = 0


@pytest.mark.parametrize("tag,Nh", TAGS)
def test_env_oracle_step_matches_reference(oracle, golden_env, golden_mpc, tag, Nh):
    g = golden_env
    _, tab, lens = mpc_path_table(golden_mpc)
    T = g[f"{tag}_out_action"].shape[0]
    env = oracle.EnvRef(T, Nh, tab, lens, auto_path_switch=False)
    _load(env, g, tag, "pre")
    a = env.step(g[f"{tag}_out_action"])
    assert np.array_equal(a["hvel"].view(np.uint32), g[f"{tag}_post_hvel"].view(np.uint32))
    assert np.abs(a["hpos"] - g[f"{tag}_post_hpos"]).max() <= 1e-12
    assert np.abs(a["hhist"] - g[f"{tag}_post_hhist"]).max() <= 1e-12
    assert np.abs(a["rpose"][:, :3] - g[f"{tag}_post_rpose"][:, :3]).max() <= 1e-12
    assert np.abs(a["rhist"][:, :, :3] - g[f"{tag}_post_rhist"][:, :, :3]).max() <= 1e-12
    assert np.array_equal(a["loc_to_start"], g[f"{tag}_post_loc_to_start"])
    changed = np.abs(g[f"{tag}_post_hgoal"] - g[f"{tag}_pre_hgoal"]).max(-1) > 0
    assert np.array_equal(a["hgoal"][~changed], g[f"{tag}_post_hgoal"][~changed])
    assert (np.abs(a["hgoal"][changed] - g[f"{tag}_pre_hgoal"][changed]).max(-1) > 0).all()
    for name, tol in (("pt_state", 1e-9), ("robot_node", 1e-9), ("past_robot_vel", 1e-15), ("percent_episode", 1e-12),
                      ("spatial_edges", 1e-9), ("human_valid", 0), ("detected_humans", 0)):
        assert np.abs(a[name] - g[f"{tag}_out_{name}"].reshape(a[name].shape)).max() <= tol, name
    assert np.abs(a["rewards"] - g[f"{tag}_out_rewards"]).max() <= 1e-9
    assert np.abs(a["info"][:, 0] - g[f"{tag}_out_disturbance_v"]).max() <= 1e-9
    assert np.abs(a["info"][:, 1] - g[f"{tag}_out_disturbance_th"]).max() <= 1e-9
    assert np.abs(a["info"][:, 2] - g[f"{tag}_out_path_advancement"]).max() <= 1e-9
    assert np.abs(a["info"][:, 3] - g[f"{tag}_out_deviation"]).max() <= 1e-9
    assert np.array_equal(a["done_bits"] & 0xFF, g[f"{tag}_out_bits"].astype(np.uint32) & 0xFF)
    d = np.abs(a["mpc_actions"] - g[f"{tag}_out_mpc_actions"]).max(1)
    print(f"{tag}: {T} transitions, MPC max|du|={d.max():.2e}, >1e-6: {int((d > 1e-6).sum())}")
    assert d.max() <= 1e-6                       # measured 1.6e-7 / 1.3e-7 / 4.8e-8: every transition, no divergence allowed


@pytest.mark.parametrize("tag,Nh", TAGS)
def test_env_oracle_reset_matches_reference(oracle, golden_env, golden_mpc, tag, Nh):
    g = golden_env
    _, tab, lens = mpc_path_table(golden_mpc)
    env = oracle.EnvRef(2, Nh, tab, lens, end_goal_changing=False)
    hp = np.broadcast_to(g[f"{tag}_reset_hpos"], (2, Nh, 2)); hg = np.broadcast_to(g[f"{tag}_reset_hgoal"], (2, Nh, 2))
    a = env.reset(hp, hg)
    assert np.array_equal(a["hvel"][0].view(np.uint32), g[f"{tag}_postreset_hvel"].view(np.uint32))
    assert np.abs(a["hpos"][0] - g[f"{tag}_postreset_hpos"]).max() <= 1e-12
    assert np.abs(a["hhist"][0] - g[f"{tag}_postreset_hhist"]).max() <= 1e-12
    for name, tol in (("pt_state", 1e-9), ("robot_node", 1e-9), ("spatial_edges", 1e-9), ("mpc_actions", 1e-6)):
        assert np.abs(a[name][0] - g[f"{tag}_reset_{name}"].reshape(a[name][0].shape)).max() <= tol, name


def test_env_oracle_spawn_and_free_run(oracle, golden_mpc):
    """Philox spawning respects the reference's rejection rules (crowd_sim.py:247-258); a free run stays finite and
    deterministic, and is independent of the thread count."""
    _, tab, lens = mpc_path_table(golden_mpc)
    outs = []
    for nth in (1, 4):
        env = oracle.EnvRef(64, 12, tab, lens)
        a = env.reset(nthreads=nth)
        if nth == 1:
            P0 = a["hhist"][:, :, 0, :]   # spawn positions are the oldest history entries
            d = np.linalg.norm(P0[:, :, None] - P0[:, None], axis=-1) + np.eye(12) * 10
            assert d.min() >= 0.8 - 1e-12
            assert np.linalg.norm(P0 - np.array([0, -4.0]), axis=-1).min() >= 1.1 - 1e-12
        for t in range(25):
            a = env.step(a["mpc_actions"][:, :2].copy(), nthreads=nth)
        assert np.isfinite(a["rewards"]).all() and np.isfinite(a["spatial_edges"]).all()
        outs.append({k: v.copy() for k, v in a.items() if k != "mpc_status"})
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
         "path_idx", "loc_to_start", "warm_T", "warm_u")


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10), ("h6s", 6)])
def test_env_oracle_soft_reset_follows_reference(oracle, golden_mpc, tag, Nh):
    """oracle/env_ref.c's restatement of HAAndPTEnv.soft_reset (HA_and_PT env.py:120-240) replays EVERY env.step call the
    reference issued in three runs (tests/golden/soft_reset.npz), the scripted ones from inside soft_reset() included: one
    env, physical state teacher-forced before every call, the loop variables and the previous done word left to evolve.
    Scripted-or-caller decision and scripted action exact; done reasons equal; rewards 1e-9; h6s has a static human."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "soft_reset.npz"))
    _, tab, lens = mpc_path_table(golden_mpc)
    T = g[f"{tag}_action"].shape[0]
    env = oracle.EnvRef(1, Nh, tab, lens, soft_reset=True)
    env.reset()
    kinds = {"caller": 0, "wait": 0, "backup": 0, "turn": 0, "creep": 0}
    for t in range(T):
        for name in STATE:
            env.a[name][...] = g[f"{tag}_pre_{name}"][t:t + 1].reshape(env.a[name].shape)
        env.a["iteration0"][...] = g[f"{tag}_pre_it0"][t:t + 1]
        soft = bool(g[f"{tag}_soft"][t])
        ref_a = g[f"{tag}_action"][t]
        a = env.step(np.array([[0.777, -0.555]]) if soft else ref_a[None], nthreads=1)   # poison: must be overridden
        assert bool(a["soft_flags"][0, 0]) == soft, (t, "scripted flag")
        assert np.array_equal(a["actions_used"][0], ref_a), (t, a["actions_used"][0], ref_a)
        assert a["soft_flags"][0, 1] == 0
        assert (int(a["done_bits"][0]) & 0xFF) == (int(g[f"{tag}_bits"][t]) & 0xFF), t
        assert np.abs(a["rewards"][0] - g[f"{tag}_rewards"][t]).max() <= 1e-9, t
        assert np.abs(a["rpose"][0, :3] - g[f"{tag}_post_rpose"][t, :3]).max() <= 1e-12
        assert np.array_equal(a["hvel"][0].view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32))
        k = ("backup" if ref_a[0] == -0.3 else "creep" if ref_a[0] == 0.5 else "turn" if ref_a[1] != 0 else "wait") if soft else "caller"
        kinds[k] += 1
    print(tag, T, "calls", kinds)
    assert kinds["caller"] > 0 and kinds["wait"] > 0
    if tag == "h6s":
        assert kinds["backup"] > 20 and g[f"{tag}_pre_hstatic"].any()


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_env_oracle_sensor_restriction_matches_reference(oracle, golden_mpc, tag, Nh):
    """Robot sensor restriction in the oracle against the reference run with the switch on (sensor_restriction.npz)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "sensor_restriction.npz"))
    _, tab, lens = mpc_path_table(golden_mpc)
    T = g[f"{tag}_out_action"].shape[0]
    env = oracle.EnvRef(T, Nh, tab, lens, auto_path_switch=False, sensor_restriction=True)
    for name in STATE:
        env.a[name][...] = g[f"{tag}_pre_{name}"].reshape(env.a[name].shape)
    env.a["iteration0"][...] = g[f"{tag}_pre_it0"]
    a = env.step(g[f"{tag}_out_action"])
    assert np.array_equal(a["hvalid"], g[f"{tag}_post_hvalid"])
    assert np.abs(a["hhist"] - g[f"{tag}_post_hhist"]).max() <= 1e-12
    assert np.array_equal(a["detected_humans"], g[f"{tag}_out_detected_humans"].reshape(-1))
    assert np.array_equal(a["human_valid"], g[f"{tag}_out_human_valid"].reshape(a["human_valid"].shape))
    assert np.abs(a["spatial_edges"] - g[f"{tag}_out_spatial_edges"].reshape(a["spatial_edges"].shape)).max() <= 1e-9
    assert np.abs(a["rewards"] - g[f"{tag}_out_rewards"]).max() <= 1e-9
    assert (a["hvalid"].sum(1) == 0).any()
    # reset(): histories of out-of-range humans start at 0 (human_avoidance_env.py:331-334)
    env2 = oracle.EnvRef(2, Nh, tab, lens, sensor_restriction=True)
    r = env2.reset(np.broadcast_to(g[f"{tag}_reset_hpos"], (2, Nh, 2)), np.broadcast_to(g[f"{tag}_reset_hgoal"], (2, Nh, 2)))
    assert r["detected_humans"][0] == g[f"{tag}_reset_detected_humans"]
    assert np.array_equal(r["human_valid"][0], g[f"{tag}_reset_human_valid"])
    assert np.abs(r["spatial_edges"][0] - g[f"{tag}_reset_spatial_edges"]).max() <= 1e-9


def test_env_oracle_static_humans(oracle, golden_mpc):
    """include_static_humans (crowd_sim.py:171-183, 270-272): the last humans are parked obstacles with a zero action."""
    _, tab, lens = mpc_path_table(golden_mpc)
    env = oracle.EnvRef(32, 8, tab, lens, static_humans=((-4.0, 0.0), (0.0, 0.0)))
    a = env.reset()
    assert (a["hstatic"][:, :6] == 0).all() and (a["hstatic"][:, 6:] == 1).all()
    for t in range(20):
        a = env.step(a["mpc_actions"][:, :2].copy())
        assert np.array_equal(a["hpos"][:, 6:], np.broadcast_to([[-4.0, 0.0], [0.0, 0.0]], (32, 2, 2))) and (a["hvel"][:, 6:] == 0).all()
