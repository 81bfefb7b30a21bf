"""GPU: the C-ABI additions of round 2 -- double-buffered output slab, masked reset / observe, the host-driven soft-reset
primitives (dre_env_switch_path, dre_env_soft_reset_decide) -- checked for the properties their callers rely on."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env(E=256, Nh=6, **kw):
    import dr_mpc_b200 as d
    return d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, **kw), num_envs=E)


def test_output_slab_is_double_buffered():
    """S of step t-1 stays intact while the caller holds S' of step t: the reference's `insert(S, a, S'); S = S'`
    (online_continuous_task.py:185-189) works on the zero-copy views; a third step reuses the first slab."""
    import torch
    import dr_mpc_b200 as d
    env = _env()
    S = env.reset()
    rb = d.BatchedReplayBuffer(S, buffer_size=4096)
    frozen = None
    for t in range(5):
        a = S['PT']['MPC_actions'][:, :2].clone()
        keep = {k: v.clone() for k, v in S['HA'].items()}
        S_prime, R, D, info, eps = env.step(a)
        torch.cuda.synchronize()
        for k, v in keep.items():
            assert torch.equal(S['HA'][k], v), (t, k)                 # S survived the step that produced S_prime
        assert S['HA']['spatial_edges'].data_ptr() != S_prime['HA']['spatial_edges'].data_ptr()
        assert not torch.equal(S['HA']['spatial_edges'], S_prime['HA']['spatial_edges'])
        rb.insert(S, a, S_prime, R['R'], D['done_for_replay'])
        if t == 0:
            frozen = S
        S = S_prime
    assert frozen['HA']['spatial_edges'].data_ptr() in (S['HA']['spatial_edges'].data_ptr(), env._outs[env._L.dre_env_current_slab(env._h) ^ 1]['spatial_edges'].data_ptr())
    with pytest.raises(ValueError, match="alias"):
        rb.insert(S, a, S, R['R'], D['done_for_replay'])
    # the views keep the engine alive
    se = S['HA']['spatial_edges']
    ref = se.clone()
    del env, S, S_prime, frozen, rb
    import gc
    gc.collect()
    torch.zeros(1 << 20, device="cuda")
    assert torch.equal(se, ref)


def test_masked_reset_leaves_the_other_envs_alone():
    import torch
    E, Nh = 300, 8
    env = _env(E, Nh)
    S = env.reset()
    for t in range(12):
        S, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    torch.cuda.synchronize()
    mask = torch.zeros(E, dtype=torch.bool, device=env.device)
    mask[::3] = True
    before = {k: v.clone() for k, v in env.state.items()}
    out_before = {k: v.clone() for k, v in env.out.items()}
    T0, u0, it0 = env.mpc_get_warm()
    S2 = env.reset(env_mask=mask)
    torch.cuda.synchronize()
    keep = ~mask
    for k, v in env.state.items():
        assert torch.equal(v[keep], before[k][keep]), k
    for k, v in env.out.items():
        assert torch.equal(v[keep], out_before[k][keep]), k
    T1, u1, it1 = env.mpc_get_warm()
    kn = keep.cpu().numpy()
    assert np.array_equal(T0[kn], T1[kn]) and np.array_equal(u0[kn], u1[kn]) and np.array_equal(it0[kn], it1[kn])
    st = env.state
    assert (st["step_count"][mask] == 0).all() and (st["gtime"][mask] == 0).all() and (st["path_idx"][mask] == 0).all()
    assert (st["rpose"][mask][:, 0] == 0).all() and (st["rpose"][mask][:, 1] == -4).all()
    assert (st["step_count"][keep] == 12).all()
    assert (st["hvalid"][mask] == 7).all() and (st["rvalid"][mask] == 7).all()        # the lookback warm-start steps ran
    # a masked reset from given spawn positions equals a full reset from them (spawning then needs no random draw; goals far
    # away so that none is reached -- and resampled from the epoch-keyed stream -- during the warm-start steps)
    hp = before["hpos"].clone(); hg = hp + torch.tensor([7.0, 3.0], dtype=torch.float64, device=hp.device)
    ref = _env(E, Nh)
    ref.reset(hp, hg)
    env.reset(hp, hg, env_mask=mask)
    torch.cuda.synchronize()
    for k in ("hpos", "hvel", "hhist", "rpose", "rhist"):
        assert torch.equal(env.state[k][mask], ref.state[k][mask]), k
    for k in ("pt_state", "mpc_actions", "spatial_edges", "robot_node"):
        assert torch.equal(env.out[k][mask], ref.out[k][mask]), k
    # and the batch keeps stepping
    S, R, D, info, eps = env.step(env.out["mpc_actions"][:, :2].clone())
    assert torch.isfinite(R['R']).all()


def test_switch_path_and_masked_observe():
    """The host-driven form of soft_reset's path switch (HA_and_PT env.py:130-148) equals what auto_path_switch does inside
    the step: same path index, localisation flag, re-armed MPC, PT state and MPC actions on the new path."""
    import torch
    E = 128
    a_env = _env(E, auto_path_switch=False)
    S = a_env.reset()
    for t in range(5):
        S, *_ = a_env.step(S['PT']['MPC_actions'][:, :2].clone())
    torch.cuda.synchronize()
    mask = torch.zeros(E, dtype=torch.bool, device=a_env.device)
    mask[1::2] = True
    pre_out = {k: v.clone() for k, v in a_env.out.items()}
    T0, u0, _ = a_env.mpc_get_warm()
    a_env.switch_path(mask)
    S2 = a_env.observe(solve_mpc=True, env_mask=mask)
    torch.cuda.synchronize()
    st = a_env.state
    assert (st["path_idx"][mask] == 1).all() and (st["path_idx"][~mask] == 0).all()
    assert (st["loc_to_start"][mask] == 1).all()
    for k, v in a_env.out.items():
        assert torch.equal(v[~mask], pre_out[k][~mask]), k
    T1, u1, it1 = a_env.mpc_get_warm()
    kn = (~mask).cpu().numpy()
    assert np.array_equal(T0[kn], T1[kn]) and np.array_equal(u0[kn], u1[kn]) and (it1 == 0).all()
    assert not torch.equal(S2['PT']['PT_state'][mask], pre_out["pt_state"][mask])
    # reference for the new path: a second engine put on path 1 by hand, first solve = iteration0
    b_env = _env(E, auto_path_switch=False)
    b_env.reset()
    for k in a_env.state:
        b_env.state[k].copy_(a_env.state[k])
    b_env.mpc_set_warm(T0, u0, np.ones(E, dtype=np.uint8))   # same stored warm start: a failing first solve retries from it
    Sb = b_env.observe(solve_mpc=True)
    torch.cuda.synchronize()
    assert torch.equal(Sb['PT']['PT_state'][mask], S2['PT']['PT_state'][mask])
    assert torch.equal(Sb['PT']['MPC_actions'][mask], S2['PT']['MPC_actions'][mask])


def test_soft_reset_decide_predicts_the_device_machine():
    """dre_env_soft_reset_decide(commit=0) before a step = what the on-device machine then does in that step: same
    scripted-or-caller flag, same scripted action; peeking leaves the loop variables alone."""
    import torch
    E = 2048
    env = _env(E, 10, soft_reset_on_device=True)
    S = env.reset()
    seen = 0
    for t in range(120):
        sr0 = env.state["sr_state"].clone()
        scripted, act, stuck = env.soft_reset_decide(commit=False)
        assert torch.equal(env.state["sr_state"], sr0)
        a = S['PT']['MPC_actions'][:, :2].clone()
        S, R, D, info, eps = env.step(a)
        assert torch.equal(info['soft_reset'], scripted), t
        assert torch.equal(info['actions_used'][scripted], act[scripted])
        assert not bool(stuck.any())
        seen += int(scripted.sum())
    assert seen > 100


def test_host_output_modes():
    """dre_env_set_host_outputs: float32 observations are the float64 results rounded once (bit-equal to numpy's cast of
    the full slab); 'control' brings back only the control fields, identical to the full slab's."""
    import torch
    import dr_mpc_b200 as d
    E, Nh = 1000 + 13, 10
    envs = {m: d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E) for m in ("full", "f32", "ctl")}
    envs["f32"].configure_host_outputs(out_dtype="float32")
    envs["ctl"].configure_host_outputs(outputs="control")
    for e in envs.values():
        e.reset()
    act = envs["full"].out["mpc_actions"][:, :2].cpu().numpy().copy()
    for t in range(5):
        v = {m: e.step_host(act) for m, e in envs.items()}
        for name in d.BatchedHAAndPTEnv.OBS_FIELDS:
            assert v["f32"][name].dtype == np.float32 and name not in v["ctl"]
            assert np.array_equal(v["f32"][name], v["full"][name].astype(np.float32)), (t, name)
        for name in d.BatchedHAAndPTEnv.CONTROL_FIELDS:
            for m in ("f32", "ctl"):
                assert np.array_equal(np.asarray(v[m][name]).view(np.uint8), np.asarray(v["full"][name]).view(np.uint8)), (t, m, name)
        # the device slab stays float64 and complete in every mode
        for m in ("f32", "ctl"):
            assert torch.equal(envs[m].out["spatial_edges"], envs["full"].out["spatial_edges"])
        act = v["full"]["mpc_actions"][:, :2].copy()
    assert envs["f32"].host_bytes_per_step < 0.65 * envs["full"].host_bytes_per_step and envs["ctl"].host_bytes_per_step < 0.15 * envs["full"].host_bytes_per_step
    assert envs["full"].d2h_ceiling_ms(reps=3) > 0
