"""CPU, build container only (needs /root/reference): the reference's OWN, UNMODIFIED env classes -- HAAndPTEnv, HAEnv,
PTEnv, Human, CrowdSim -- run with the drop-in adapters of dr_mpc_b200.compat plugged in where the reference plugs in
Python-RVO2 and pysteam: policy_factory['orca'] = compat.ORCA (scripts/policy/policy_factory.py:8) and env.MPC =
compat.MPC(...) (HA_and_PT env.py:47). There is no GPU here, so the two C-ABI calls the adapters make are answered by the
CPU oracle INSIDE THIS TEST (monkeypatched; the product never imports the oracle): what is tested is everything around
those calls -- JointState packing, agent order, radii, preferred velocity, path registration, warm-start ownership,
return types -- by reproducing tests/golden/env_rollouts.npz, which the same classes produced on the rvo2 / pysteam shims.
The GPU twin (tests/test_gpu_compat.py) runs the same adapters over the real C ABI against the golden boundary traffic."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO, mpc_path_table

REF = os.environ.get("DRE_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "environment")), reason="the reference tree is only in the build container")


@pytest.fixture()
def reference_with_adapters(oracle, monkeypatch):
    for p in (os.path.join(REPO, "oracle", "shims"), REF):
        if p not in sys.path:
            monkeypatch.syspath_prepend(p)
    import dr_mpc_b200 as d
    from dr_mpc_b200 import compat
    import dr_mpc_b200.mpc as mpc_mod
    calls = {"orca": 0, "mpc": 0}

    def orca_backend(self, params, self8, others5):        # stands in for dre_orca_solve_host
        out = np.empty((self8.shape[0], 2), dtype=np.float32)
        for n in range(self8.shape[0]):
            out[n], _ = oracle.orca_solve(self8[n], others5[n], neighbor_dist=params.neighbor_dist, max_neighbors=params.max_neighbors,
                                          time_horizon=params.time_horizon, time_step=params.time_step)
            calls["orca"] += 1
        return out

    class OracleBatchedMPC:                                 # stands in for BatchedMPC (dre_mpc_create / set_paths / act_host)
        def __init__(self, config_master, K, time_step, continuous_task, max_v, max_w, num_problems=1):
            self.K, self.N, self.mb, self.last_status = K, num_problems, None, None
            self.prm = oracle.mpc_params(K=K, dt=time_step, vmax=max_v, wmax=max_w, node_separation=config_master.config_PT.env.node_separation)

        def set_paths(self, paths):
            tab, lens = d.paths.pack_paths(paths)
            old = self.mb
            self.mb = oracle.MpcBatch(self.N, self.prm, tab, lens)
            if old is not None:
                self.mb.warm_T[:], self.mb.warm_u[:], self.mb.iteration0[:] = old.warm_T, old.warm_u, old.iteration0

        def reset(self, mask=None):
            if self.mb is not None:
                self.mb.reset()

        def act(self, inp, want_status=False):
            calls["mpc"] += 1
            u, st = self.mb.act(inp['path_id'], inp['interp_index'], inp['pose'])
            self.last_status = st
            return u

    monkeypatch.setattr(compat.ORCA, "_solve", orca_backend)
    monkeypatch.setattr(mpc_mod, "BatchedMPC", OracleBatchedMPC)
    from scripts.policy import policy_factory as PF
    monkeypatch.setitem(PF.policy_factory, 'orca', compat.ORCA)
    return compat, calls


def _config(num_humans):
    from scripts.configs import ConfigMaster
    cm = ConfigMaster()
    cm.config_HA.sim.human_num = num_humans
    cm.config_HA.sim.max_allowable_humans = num_humans
    return cm


_BITS = {'collision': 1 << 0, 'safety_human_raise': 1 << 1, 'actuation_termination': 1 << 2, 'timeout': 1 << 3,
         'goal': 1 << 4, 'deviation_end_of_path': 1 << 5, 'corridor_hit': 1 << 6, 'safety_corridor_raise': 1 << 7}


@pytest.mark.parametrize("tag,Nh,seed", [("h6", 6, 0), ("h10", 10, 1), ("h20", 20, 2)])
def test_reference_env_with_adapters_reproduces_golden_rollout(reference_with_adapters, golden_env, tag, Nh, seed):
    compat, calls = reference_with_adapters
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot, ActionXY
    g = golden_env
    cm = _config(Nh)
    env = HAAndPTEnv(cm, seed_addition=seed)
    assert type(env.HA_env).__module__.startswith("environment.")
    # the second plug: the MPC instance the env hands to PTEnv (HA_and_PT env.py:44-47)
    env.MPC = compat.MPC(cm, cm.config_PT.MPC.planning_horizon, env.time_step, env.continuous_task, cm.config_general.robot.v_max, cm.config_general.robot.w_max)
    env.case_counter['train'] = seed
    S = env.reset()
    assert isinstance(env.HA_env.humans[0].policy, compat.ORCA)
    for k, name in (("pt_state", ('PT', 'PT_state')), ("spatial_edges", ('HA', 'spatial_edges')), ("robot_node", ('HA', 'robot_node'))):
        assert np.abs(S[name[0]][name[1]][0] - g[f"{tag}_reset_{k}"]).max() <= 1e-9, k
    assert np.abs(S['PT']['MPC_actions'][0] - g[f"{tag}_reset_mpc_actions"]).max() <= 1e-6
    T = g[f"{tag}_out_action"].shape[0]
    done_return, info = {'done': False}, None
    worst = {"mpc": 0.0, "obs": 0.0, "rew": 0.0, "drift": 0.0}
    tight_warm = 0
    for t in range(T):
        if done_return['done']:
            S = env.soft_reset(done_return, info, S)            # the reference's own loop, stepping through the adapters
            assert S is not None
        a = g[f"{tag}_out_action"][t]
        # The warm start lives inside the adapter. Its iteration0 flag must have followed the reference's resets on its own;
        # the stored poses / controls are then aligned with the recording before the solve (teacher forcing, like every
        # other MPC parity test here): the reference's LM is chaotic in its own warm start while the controls sit on their
        # bounds after a path switch (a 4e-8 difference becomes 9e-5 one solve later, measured), so a free-running
        # comparison cannot hold 1e-6 there. `drift` records how far the adapter's own warm start had moved.
        mb = env.MPC._mpc.mb
        assert int(mb.iteration0[0]) == int(g[f"{tag}_pre_it0"][t]), t
        if not g[f"{tag}_pre_it0"][t]:
            worst["drift"] = max(worst["drift"], np.abs(mb.warm_T[0] - g[f"{tag}_pre_warm_T"][t]).max())
            tight_warm += int(np.abs(mb.warm_T[0] - g[f"{tag}_pre_warm_T"][t]).max() <= 1e-6)
            mb.warm_T[0], mb.warm_u[0] = g[f"{tag}_pre_warm_T"][t], g[f"{tag}_pre_warm_u"][t]
        S, R, done_return, info, eps = env.step(ActionRot(float(a[0]), float(a[1])))
        assert np.abs(mb.warm_T[0] - g[f"{tag}_post_warm_T"][t]).max() <= 1e-6 and np.abs(mb.warm_u[0] - g[f"{tag}_post_warm_u"][t]).max() <= 1e-6, t
        # human velocities are what compat.ORCA returned through Human.act -> bit-exact float32 values
        hv = np.array([[h.vx, h.vy] for h in env.HA_env.humans], dtype=np.float32)
        assert np.array_equal(hv.view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32)), t
        hp = np.array([[h.px, h.py] for h in env.HA_env.humans])
        assert np.abs(hp - g[f"{tag}_post_hpos"][t]).max() <= 1e-12, t
        bits = sum(b for name, b in _BITS.items() if name in info['done'])
        assert bits == int(g[f"{tag}_out_bits"][t]), t
        worst["rew"] = max(worst["rew"], np.abs(np.array([R['R_PT'], R['R_DO'], R['R']]) - g[f"{tag}_out_rewards"][t]).max())
        worst["obs"] = max(worst["obs"], np.abs(S['HA']['spatial_edges'][0] - g[f"{tag}_out_spatial_edges"][t]).max(),
                           np.abs(S['PT']['PT_state'][0] - g[f"{tag}_out_pt_state"][t]).max())
        worst["mpc"] = max(worst["mpc"], np.abs(S['PT']['MPC_actions'][0] - g[f"{tag}_out_mpc_actions"][t]).max())
    print(f"{tag}: {T} steps through the reference's HAAndPTEnv with compat.ORCA / compat.MPC: max |dMPC| {worst['mpc']:.2e}, "
          f"|dobs| {worst['obs']:.2e}, |dR| {worst['rew']:.2e}; free-running warm start within 1e-6 of the recording on {tight_warm}/{T} steps "
          f"(max drift {worst['drift']:.1e}); adapter calls {calls}")
    assert worst["mpc"] <= 1e-6 and worst["obs"] <= 1e-9 and worst["rew"] <= 1e-9
    assert calls["orca"] >= 2 * Nh * T and calls["mpc"] >= T


def test_adapter_types_and_state_packing(reference_with_adapters):
    """compat.ORCA.predict takes the reference's JointState and returns the reference's ActionXY; the packed arrays are what
    orca.py:84-111 would hand to rvo2 (ego first, then human_states in order; radius + 0.01 + safety_space; v_pref as the
    ego's max speed; preferred velocity un-normalised within 1 m of the goal)."""
    compat, calls = reference_with_adapters
    from environment.human_avoidance.utils.state import JointState
    from environment.human_avoidance.utils.action import ActionXY
    cm = _config(6)
    pol = compat.ORCA(cm.config_HA)
    pol.time_step = 0.25
    js = JointState((0.0, 0.0, 0.1, 0.0, 0.3, 0.5, 0.2, 1.0, 0.0), [(1.0, 0.0, 0.0, 0.0, 0.3), (0.0, 2.0, -0.5, 0.0, 0.4)])
    self8, others = compat.orca_pack_joint_state(js, pol.safety_space)
    assert np.allclose(self8, [0, 0, 0.1, 0, 0.5, 0.2, 0.485, 1.0]) and self8.dtype == np.float32
    assert np.allclose(others, [[1, 0, 0, 0, 0.485], [0, 2, -0.5, 0, 0.585]])
    a = pol.predict(js)
    assert isinstance(a, ActionXY) and pol.last_state is js and pol.max_neighbors == 2
    far = JointState((0.0, 0.0, 0.0, 0.0, 0.3, 3.0, 4.0, 1.0, 0.0), [(5.0, 5.0, 0.0, 0.0, 0.3)])
    s8, _ = compat.orca_pack_joint_state(far, 0.175)
    assert np.allclose(s8[4:6], [0.6, 0.8])
    many = pol.predict_many([js, js])
    assert many[0] == a and many[1] == a


@pytest.fixture()
def reference_on_emulated_library(monkeypatch):
    """The same plugs, but nothing stands in for the C ABI: dr_mpc_b200._lib loads the EMULATED build of the product library
    (tests/host_harness/gen_emu.py: every .cu compiled for the CPU, CUDA threads as coroutines), so compat.ORCA's
    dre_orca_solve_host and compat.MPC's dre_mpc_create / set_paths / reset / act_host run the product's own kernels."""
    for p in (os.path.join(REPO, "oracle", "shims"), REF):
        if p not in sys.path:
            monkeypatch.syspath_prepend(p)
    import emu_driver
    import dr_mpc_b200._lib as LL
    emu = emu_driver.lib()
    monkeypatch.setattr(LL, "_lib", emu)
    from dr_mpc_b200 import compat
    from scripts.policy import policy_factory as PF
    monkeypatch.setitem(PF.policy_factory, 'orca', compat.ORCA)
    return compat


@pytest.mark.parametrize("tag,Nh,seed,steps", [("h6", 6, 0, 500), ("h10", 10, 1, 90), ("h20", 20, 2, 50)])
def test_reference_env_runs_on_the_products_kernels(reference_on_emulated_library, golden_env, tag, Nh, seed, steps):
    """The reference's UNMODIFIED HAAndPTEnv / HAEnv / PTEnv / Human / CrowdSim, with policy_factory['orca'] = compat.ORCA and
    env.MPC = compat.MPC, stepping on the product's own orca_joint_kernel and mpc_coop_kernel (emulated GPU): the recorded
    rollout -- resets, path switches and the reference's own soft_reset() loop included -- is reproduced: human velocities
    bit-exact, positions 1e-12, observations / rewards 1e-9, done reasons equal, MPC controls 1e-6."""
    compat = reference_on_emulated_library
    from environment.HA_and_PT.human_avoidance_and_path_tracking_env import HAAndPTEnv
    from environment.human_avoidance.utils.action import ActionRot
    g = golden_env
    cm = _config(Nh)
    env = HAAndPTEnv(cm, seed_addition=seed)
    env.MPC = compat.MPC(cm, cm.config_PT.MPC.planning_horizon, env.time_step, env.continuous_task, cm.config_general.robot.v_max, cm.config_general.robot.w_max)
    env.case_counter['train'] = seed
    S = env.reset()
    assert isinstance(env.HA_env.humans[0].policy, compat.ORCA)
    assert np.abs(S['HA']['spatial_edges'][0] - g[f"{tag}_reset_spatial_edges"]).max() <= 1e-9
    assert np.abs(S['PT']['MPC_actions'][0] - g[f"{tag}_reset_mpc_actions"]).max() <= 1e-6
    T = min(steps, g[f"{tag}_out_action"].shape[0])
    done_return, info = {'done': False}, None
    worst = {"mpc": 0.0, "obs": 0.0, "rew": 0.0}
    mpc = env.MPC._mpc                          # the product's BatchedMPC over the emulated C ABI
    for t in range(T):
        if done_return['done']:
            S = env.soft_reset(done_return, info, S)
            assert S is not None
        a = g[f"{tag}_out_action"][t]
        wT, wu, it0 = mpc.get_warm()
        assert int(it0[0]) == int(g[f"{tag}_pre_it0"][t]), t            # the adapter followed the reference's MPC.reset() calls
        if not g[f"{tag}_pre_it0"][t]:                                   # teacher forcing of the warm start, as in every MPC parity test
            mpc.set_warm(g[f"{tag}_pre_warm_T"][t][None], g[f"{tag}_pre_warm_u"][t][None], None)
        S, R, done_return, info, eps = env.step(ActionRot(float(a[0]), float(a[1])))
        hv = np.array([[h.vx, h.vy] for h in env.HA_env.humans], dtype=np.float32)
        assert np.array_equal(hv.view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32)), t
        hp = np.array([[h.px, h.py] for h in env.HA_env.humans])
        assert np.abs(hp - g[f"{tag}_post_hpos"][t]).max() <= 1e-12, t
        assert sum(b for name, b in _BITS.items() if name in info['done']) == int(g[f"{tag}_out_bits"][t]), t
        worst["rew"] = max(worst["rew"], np.abs(np.array([R['R_PT'], R['R_DO'], R['R']]) - g[f"{tag}_out_rewards"][t]).max())
        worst["obs"] = max(worst["obs"], np.abs(S['HA']['spatial_edges'][0] - g[f"{tag}_out_spatial_edges"][t]).max(),
                           np.abs(S['PT']['PT_state'][0] - g[f"{tag}_out_pt_state"][t]).max())
        worst["mpc"] = max(worst["mpc"], np.abs(S['PT']['MPC_actions'][0] - g[f"{tag}_out_mpc_actions"][t]).max())
    print(f"{tag}: {T} steps of the reference's env on the emulated product kernels: max |dMPC| {worst['mpc']:.2e}, |dobs| {worst['obs']:.2e}, |dR| {worst['rew']:.2e}")
    assert worst["mpc"] <= 1e-6 and worst["obs"] <= 1e-9 and worst["rew"] <= 1e-9
