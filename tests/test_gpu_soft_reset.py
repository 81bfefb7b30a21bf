"""GPU parity of the on-device soft-reset state machine against the reference's own HAAndPTEnv.soft_reset
(HA_and_PT env.py:120-240), run unmodified on the shims (tests/golden/soft_reset.npz holds EVERY env.step call of three
runs, including the scripted calls issued from inside soft_reset(), with the action the reference chose).

The device env is stepped once per recorded call. The physical state is teacher-forced from the recording before every
step (so MPC round-off cannot drift the trajectories apart); the soft-reset machine's own state is NOT: it lives on the
device and has to evolve, from the device's own done bits, through the same waits / back-ups / turns as the reference.
Checked per call: scripted-or-caller flag, the executed action (exact), done reasons, rewards (1e-9), post-step poses."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "soft_reset.npz")
STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
         "path_idx", "loc_to_start")


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10), ("h6s", 6)])
def test_soft_reset_machine_follows_reference(tag, Nh):
    import torch
    import dr_mpc_b200 as d
    g = np.load(GOLDEN)
    T = g[f"{tag}_action"].shape[0]
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, soft_reset_on_device=True), num_envs=1)
    env.reset()
    dev = env.device
    kinds = {"caller": 0, "wait": 0, "backup": 0, "turn": 0, "creep": 0}
    bad_mpc = 0
    for t in range(T):
        for name in STATE:
            a = np.ascontiguousarray(g[f"{tag}_pre_{name}"][t:t + 1])
            env.state[name].copy_(torch.as_tensor(a, device=dev).reshape(env.state[name].shape))
        env.mpc_set_warm(g[f"{tag}_pre_warm_T"][t:t + 1], g[f"{tag}_pre_warm_u"][t:t + 1], g[f"{tag}_pre_it0"][t:t + 1].astype(np.uint8))
        soft = bool(g[f"{tag}_soft"][t])
        ref_a = g[f"{tag}_action"][t]
        # a scripted call gets a poison action from the caller: the engine must override it
        a_in = torch.tensor(np.array([[0.777, -0.555]] if soft else [ref_a]), dtype=torch.float64, device=dev)
        S, R, D, info, eps = env.step(a_in)
        torch.cuda.synchronize()
        assert bool(info['soft_reset'][0]) == soft, (t, "scripted flag")
        used = info['actions_used'][0].cpu().numpy()
        assert np.array_equal(used, ref_a), (t, used, ref_a)
        assert not bool(info['soft_reset_stuck'][0])
        bits = int(info['done_bits'][0])
        assert (bits & 0xFF) == (int(g[f"{tag}_bits"][t]) & 0xFF), (t, bits, int(g[f"{tag}_bits"][t]))
        assert bool(D['done'][0]) == bool(g[f"{tag}_done"][t]) and bool(D['done_HA'][0]) == bool(g[f"{tag}_done_HA"][t])
        rw = env.out["rewards"][0].cpu().numpy()
        assert np.abs(rw - g[f"{tag}_rewards"][t]).max() <= 1e-9, (t, rw, g[f"{tag}_rewards"][t])
        assert np.abs(env.state["rpose"][0, :3].cpu().numpy() - g[f"{tag}_post_rpose"][t, :3]).max() <= 1e-12
        assert np.array_equal(env.state["hvel"][0].cpu().numpy().view(np.uint32), g[f"{tag}_post_hvel"][t].view(np.uint32))
        if soft:
            k = "backup" if ref_a[0] == -0.3 else "creep" if ref_a[0] == 0.5 else "turn" if ref_a[1] != 0 else "wait"
        else:
            k = "caller"
        kinds[k] += 1
    print(tag, T, "calls", kinds)
    assert kinds["caller"] > 0 and kinds["wait"] > 0
    if tag == "h6":
        assert kinds["turn"] > 0 and kinds["creep"] > 0
    if tag == "h6s":
        assert kinds["backup"] > 20


def test_soft_reset_off_leaves_actions_alone():
    import torch
    import dr_mpc_b200 as d
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=6), num_envs=256)
    S = env.reset()
    for _ in range(40):
        a = S['PT']['MPC_actions'][:, :2].clone()
        S, R, D, info, eps = env.step(a)
        assert torch.equal(info['actions_used'], a) and not bool(info['soft_reset'].any())


def test_soft_reset_free_run_recovers_every_env():
    """Free-running batch with the machine on: scripted transitions only follow terminations, every env gets back to
    caller control, and nothing is reported stuck."""
    import torch
    import dr_mpc_b200 as d
    E = 4096
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=10, soft_reset_on_device=True), num_envs=E)
    S = env.reset()
    prev_done = torch.zeros(E, dtype=torch.bool, device=env.device)
    in_sr = torch.zeros(E, dtype=torch.bool, device=env.device)
    scripted = 0
    left = torch.zeros(E, dtype=torch.bool, device=env.device)
    for t in range(150):
        a = S['PT']['MPC_actions'][:, :2].clone()
        S, R, D, info, eps = env.step(a)
        soft = info['soft_reset'].clone()
        # a scripted step can only start right after a done step, and continues only while the machine is active
        assert not bool((soft & ~in_sr & ~prev_done).any())
        left |= in_sr & ~soft
        scripted += int(soft.sum())
        in_sr = soft
        prev_done = D['done'].clone()
        assert not bool(info['soft_reset_stuck'].any())
        assert torch.equal(info['actions_used'][~soft], a[~soft])
    assert scripted > 0 and bool(left.any())
    assert torch.isfinite(env.out["rewards"]).all()
