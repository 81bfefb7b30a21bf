"""GPU: the literal drop-in adapters (dr_mpc_b200.compat: the reference's own call signatures, E = 1, numpy in / out) over
the real C ABI, against the golden traffic recorded at those very call sites from the reference's unmodified Python
(oracle/make_golden.py): ORCA.predict(JointState), MPC.act(dict), HAAndPTEnv.step / soft_reset /
cvmm_diverse_safety_pipeline. tests/test_compat_reference.py is the CPU twin that drives the reference's own env classes
through the same adapters."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ORCA_KEYS, mpc_path_table

pytestmark = pytest.mark.gpu


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _config_master(num_humans=6):
    """The attribute bags the adapters read, with the reference's default values (scripts/configs/*.py)."""
    import dr_mpc_b200 as d
    c = d.EngineConfig(num_humans=num_humans)
    return _Bag(
        config_general=_Bag(env=_Bag(lookback=6, time_step=0.25, time_limit=50, continuous_task=True), robot=_Bag(v_max=1.0, w_max=1.0),
                            model=_Bag(policy='DR-MPC', use_PEB=True), device='cpu', seed=0),
        config_HA=_Bag(sim=_Bag(human_num=num_humans, circle_radius=4, circle_radius_humans=5),
                       orca=_Bag(neighbor_dist=10, time_horizon=5, time_horizon_obst=5, safety_space=0.175),
                       humans=_Bag(radius=0.3, v_max=1.0, end_goal_changing=True), robot=_Bag(radius=0.3, visible=True),
                       rewards=_Bag(params={'actuation_termination': {'min_vel': 0.025, 'penalty': -20}, 'safety_human_raise': {'safety_dist': 0.22, 'penalty': -15},
                                            'collision': {'penalty': -15}, 'disturbance': {'radius': 1.5, 'factor_v': 0.8 * 7, 'factor_th': 0.5 * 7}})),
        config_PT=_Bag(env=_Bag(node_separation=0.2), MPC=_Bag(planning_horizon=4, actions_in_input=4), PT_model_params={'MLP_local': _Bag(lookahead=20, lookbehind=5)},
                       rewards=_Bag(overall_scaling=0.5, params={'goal': {'goal_radius': 0.3, 'goal_angle': 0.5, 'goal_reward': 0}, 'path_advancement': {'path_advancement_scaling': 2.0},
                                                                 'deviation': {'xy': 1, 'th': 0.05, 'scale': 0.9}, 'corridor_hit': {'penalty': -20, 'max_deviation': 1.6},
                                                                 'deviation_end_of_path': {'penalty': -10}, 'safety_corridor_raise': {'penalty': -20, 'max_deviation': 1.4}})))


def test_orca_predict_joint_state_bit_exact(golden_orca):
    """ORCA(config).predict(JointState) -> ActionXY for every human of every golden crowd, states built exactly like
    CrowdSim.get_human_actions / Human.act do (crowd_sim.py:274-280, human.py:23: other humans in list order, robot last
    and at rest): the float32 velocities of the reference, bit for bit."""
    from dr_mpc_b200 import compat
    cm = _config_master()
    pol = compat.ORCA(cm.config_HA)
    pol.time_step = 0.25
    n_single = n_all = 0
    for key in ORCA_KEYS:
        hpos, hvel, goal, rpos, vnew = (golden_orca[f"{key}_{k}"] for k in ("hpos", "hvel", "goal", "rpos", "vnew"))
        C, Nh, _ = hpos.shape
        for c in range(C):
            states = []
            for i in range(Nh):
                me = compat.FullState(hpos[c, i, 0], hpos[c, i, 1], float(hvel[c, i, 0]), float(hvel[c, i, 1]), 0.3, goal[c, i, 0], goal[c, i, 1], 1.0, 0.0)
                ob = [compat.ObservableState(hpos[c, j, 0], hpos[c, j, 1], float(hvel[c, j, 0]), float(hvel[c, j, 1]), 0.3) for j in range(Nh) if j != i]
                ob.append(compat.ObservableState(rpos[c, 0], rpos[c, 1], 0.0, 0.0, 0.3))
                states.append(_Bag(self_state=me, human_states=ob))
            acts = pol.predict_many(states)
            got = np.array([[a.vx, a.vy] for a in acts], dtype=np.float32)
            assert np.array_equal(got.view(np.uint32), vnew[c].view(np.uint32)), (key, c)
            n_all += Nh
            if c < 3:                                    # and through the one-state call the reference makes
                for i in (0, Nh - 1):
                    a = pol.predict(states[i])
                    assert isinstance(a, compat.ActionXY) and pol.last_state is states[i]
                    assert np.array_equal(np.array([a.vx, a.vy], dtype=np.float32).view(np.uint32), vnew[c, i].view(np.uint32))
                    n_single += 1
    print(f"{n_all} joint states through predict_many, {n_single} through predict: bit-exact")


@pytest.mark.parametrize("K", [4, 10, 20])
def test_mpc_act_follows_golden_sequences(golden_mpc, K):
    """MPC(config_master, K, dt, True, v_max, w_max).act({'path','interp_index','pose'}) along the four recorded sequences,
    one adapter per path like the recording (fresh object, mid-sequence reset()). The adapter owns the warm start: its
    iteration0 flag must follow on its own; the stored guess is aligned with the recording before each warm solve (the
    reference's LM is chaotic in its own warm start, see tests/test_compat_reference.py) and compared after it."""
    from dr_mpc_b200 import compat
    g = golden_mpc
    paths, tab, lens = mpc_path_table(g)
    cm = _config_master()
    pid = g[f"K{K}_path_id"]
    worst = drift = 0.0
    for p in range(4):
        idx = np.nonzero(pid == p)[0]
        m = compat.MPC(cm, K, 0.25, True, 1.0, 1.0)
        for n, t in enumerate(idx):
            if n == len(idx) // 2:
                m.reset()
            T, u, it0 = m._mpc.get_warm()
            assert int(it0[0]) == int(g[f"K{K}_it0"][t]) and m.iteration0 == bool(g[f"K{K}_it0"][t]), (p, n)
            if not g[f"K{K}_it0"][t]:
                drift = max(drift, np.abs(T[0] - g[f"K{K}_warm_T"][t]).max())
                m._mpc.set_warm(g[f"K{K}_warm_T"][t][None], g[f"K{K}_warm_u"][t][None], None)
            out = m.act({'path': paths[p], 'interp_index': g[f"K{K}_interp"][t], 'pose': g[f"K{K}_pose"][t]})
            assert out.shape == (2 * K,) and out.dtype == np.float64
            worst = max(worst, np.abs(out - g[f"K{K}_u"][t]).max())
            assert int(m.last_status["iterations"]) == int(g[f"K{K}_iters"][t]) and int(m.last_status["retries"]) == int(g[f"K{K}_retries"][t])
    print(f"K={K}: {len(pid)} act() calls, max|du| {worst:.2e}, one-solve warm-start drift {drift:.1e}")
    assert worst <= 1e-6


STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime")


def _upload(env, g, tag, t, with_path=True):
    import torch
    e = env._env
    names = STATE + (("path_idx", "loc_to_start") if with_path else ())
    for name in names:
        a = np.ascontiguousarray(g[f"{tag}_pre_{name}"][t:t + 1])
        e.state[name].copy_(torch.as_tensor(a, device=e.device).reshape(e.state[name].shape))
    e.mpc_set_warm(g[f"{tag}_pre_warm_T"][t:t + 1], g[f"{tag}_pre_warm_u"][t:t + 1], g[f"{tag}_pre_it0"][t:t + 1].astype(np.uint8) if with_path else None)
    torch.cuda.synchronize()


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10), ("h6s", 6)])
def test_env_adapter_runs_the_reference_loop(tag, Nh):
    """compat.HAAndPTEnv driven by the reference's data-collection loop (online_continuous_task.py:139-189: `if done: S =
    env.soft_reset(full_done, info, S)` ... `env.step(ActionRot, model_info=...)`) against the recording of that loop on the
    reference (tests/golden/soft_reset.npz: EVERY env.step call, the scripted ones issued from inside soft_reset() included).
    The physical state is teacher-forced before every call; which calls happen, in which order, with which action
    (scripted or the caller's), the path index / localisation flag / MPC re-arm across soft_reset's path switches, done
    reasons and rewards must all come out of the adapter."""
    import torch
    from dr_mpc_b200 import compat
    g = np.load(os.path.join(GOLDEN, "soft_reset.npz"))
    T = g[f"{tag}_action"].shape[0]
    env = compat.HAAndPTEnv(_config_master(Nh))
    S = env.reset()
    for k, shape in (('PT_state', (1, 100)), ('MPC_actions', (1, 8))):
        assert isinstance(S['PT'][k], np.ndarray) and S['PT'][k].shape == shape and S['PT'][k].dtype == np.float64
    assert S['HA']['spatial_edges'].shape == (1, Nh, 14) and S['HA']['robot_node'].shape == (1, 12) and S['HA']['detected_human_num'].shape == (1,)
    cursor = {"t": 0, "soft": 0, "switches": 0}
    inner = env.step

    def recorded_step(action, update=True, soft_reset=False, model_info=None):
        t = cursor["t"]
        assert t < T, "more env.step calls than the reference made"
        assert bool(soft_reset) == bool(g[f"{tag}_soft"][t]), (t, "scripted-or-caller")
        assert np.array_equal(np.array([action.v, action.w]), g[f"{tag}_action"][t]), (t, action, g[f"{tag}_action"][t])
        e = env._env
        # what soft_reset's path switch must have left behind, BEFORE the physical state is aligned with the recording
        assert int(e.state["path_idx"][0]) == int(g[f"{tag}_pre_path_idx"][t]), (t, "path index")
        assert int(e.state["loc_to_start"][0]) == int(g[f"{tag}_pre_loc_to_start"][t]), (t, "localize_to_start")
        assert int(e.mpc_get_warm()[2][0]) == int(g[f"{tag}_pre_it0"][t]), (t, "iteration0")
        _upload(env, g, tag, t, with_path=False)
        r = inner(action, update=update, soft_reset=soft_reset, model_info=model_info)
        S2, R, D, info, eps = r
        bits = int(g[f"{tag}_bits"][t])
        names = {n for n, b in compat._REASONS if bits & b}
        assert info['done'] == {'False'} | names, (t, info['done'], names)
        assert D['done'] == bool(g[f"{tag}_done"][t]) and D['done_HA'] == bool(g[f"{tag}_done_HA"][t])
        assert np.abs(np.array([R['R_PT'], R['R_DO'], R['R']]) - g[f"{tag}_rewards"][t]).max() <= 1e-9, t
        assert np.abs(e.state["rpose"][0, :3].cpu().numpy() - g[f"{tag}_post_rpose"][t, :3]).max() <= 1e-12
        assert np.array_equal(e.state["hvel"][0].cpu().numpy().view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32))
        assert type(R['R']) is float and type(D['done']) is bool and isinstance(S2['HA']['spatial_edges'], np.ndarray)
        cursor["t"] += 1
        cursor["soft"] += int(soft_reset)
        return r

    env.step = recorded_step
    full_done, info = {'done': False}, None
    while cursor["t"] < T:
        if full_done['done']:
            p0 = int(env._env.state["path_idx"][0])
            S = env.soft_reset(full_done, info, S)
            assert S is not None
            cursor["switches"] += int(int(env._env.state["path_idx"][0]) != p0)
            if cursor["t"] >= T:
                break
        t = cursor["t"]
        assert not g[f"{tag}_soft"][t], (t, "the reference was still inside soft_reset() here")
        a = g[f"{tag}_action"][t]
        S, full_reward, full_done, info, done_info = env.step(compat.ActionRot(float(a[0]), float(a[1])), model_info={'ID': True})
    print(f"{tag}: {T} env.step calls replayed through the adapter, {cursor['soft']} issued by soft_reset(), {cursor['switches']} path switches")
    assert cursor["soft"] == int(g[f"{tag}_soft"].sum()) and cursor["soft"] > 0
    if tag != "h6s":
        assert cursor["switches"] > 0


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_env_adapter_cvmm_matches_reference(tag, Nh):
    """cvmm_safety_check(np(2,), steps) -> (bool, dist) and cvmm_diverse_safety_pipeline(np(2,), steps, action_set) ->
    (ActionRot, info) on the recorded queries (tests/golden/cvmm.npz)."""
    import torch
    from dr_mpc_b200 import compat
    g = np.load(os.path.join(GOLDEN, "cvmm.npz"))
    aset, nsteps = g["action_set"], int(g["lookahead_steps"])
    Q = g[f"{tag}_action"].shape[0]
    env = compat.HAAndPTEnv(_config_master(Nh))
    env.reset()
    checked = 0
    for q in range(0, Q, 3):
        _upload(env, g, tag, q)
        a = g[f"{tag}_action"][q]
        safe, dist = env.cvmm_safety_check(a, nsteps)
        assert safe == bool(g[f"{tag}_cand_safe"][q, 0])
        if safe:
            assert abs(dist - g[f"{tag}_cand_dist"][q, 0]) <= 1e-9
        act, info = env.cvmm_diverse_safety_pipeline(a, nsteps, aset)
        assert isinstance(act, compat.ActionRot) and isinstance(info['RL_action_bool'], str)
        if not g[f"{tag}_random"][q]:
            assert np.array_equal(np.array([act.v, act.w]), g[f"{tag}_safe_action"][q]), q
            assert info['RL_action_bool'].startswith('NA') == bool(g[f"{tag}_cand_safe"][q, 0])
        else:
            assert info['RL_action_bool'].startswith('no safe') and any(np.array_equal([act.v, act.w], b) for b in aset)
        checked += 1
    print(tag, checked, "queries")
