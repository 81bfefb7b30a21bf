"""GPU parity of the robot sensor restriction (config_HA.robot.sensor_restriction, off by default upstream) against the
reference's own HAAndPTEnv.step / reset run with the switch on (tests/golden/sensor_restriction.npz, produced by
oracle/make_golden.py from the unmodified reference on the shims): humans beyond sensor_range lose their history
(human_avoidance_env.py:102-105,161-164,331-334) and the observation lists the visible ones first, with the fake far human
when nobody is in range (:53-72). Every recorded transition is one env of a batch (teacher forcing)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sensor_restriction.npz")
STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
         "path_idx", "loc_to_start")


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_step_with_sensor_restriction_matches_reference(tag, Nh):
    import torch
    import dr_mpc_b200 as d
    g = np.load(GOLDEN)
    T = g[f"{tag}_out_action"].shape[0]
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, robot_sensor_restriction=True, auto_path_switch=False), num_envs=T)
    env.reset()
    for name in STATE:
        a = np.ascontiguousarray(g[f"{tag}_pre_{name}"])
        env.state[name].copy_(torch.as_tensor(a, device=env.device).reshape(env.state[name].shape))
    env.mpc_set_warm(g[f"{tag}_pre_warm_T"], g[f"{tag}_pre_warm_u"], g[f"{tag}_pre_it0"].astype(np.uint8))
    env.step(torch.as_tensor(g[f"{tag}_out_action"], device=env.device))
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in env.out.items()}
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert np.array_equal(st["hvalid"], g[f"{tag}_post_hvalid"])
    assert np.abs(st["hhist"] - g[f"{tag}_post_hhist"]).max() <= 1e-12
    assert np.array_equal(o["detected_humans"], g[f"{tag}_out_detected_humans"].reshape(-1))
    assert np.array_equal(o["human_valid"], g[f"{tag}_out_human_valid"].reshape(o["human_valid"].shape))
    ref = g[f"{tag}_out_spatial_edges"].reshape(o["spatial_edges"].shape)
    assert np.abs(o["spatial_edges"] - ref).max() <= 1e-9
    assert np.abs(o["rewards"] - g[f"{tag}_out_rewards"]).max() <= 1e-9
    det = o["detected_humans"]
    hv = st["hvalid"]
    print(tag, T, "transitions: visible humans", det.min(), "..", det.max(), "| nobody in range:", int((hv.sum(1) == 0).sum()),
          "| dropped this step:", int(((g[f'{tag}_pre_hvalid'] > 0) & (hv == 0)).sum()))
    assert (hv.sum(1) == 0).any() and ((g[f"{tag}_pre_hvalid"] > 0) & (hv == 0)).any() and ((g[f"{tag}_pre_hvalid"] == 0) & (hv == 1)).any()


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_reset_with_sensor_restriction_matches_reference(tag, Nh):
    import torch
    import dr_mpc_b200 as d
    g = np.load(GOLDEN)
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, robot_sensor_restriction=True), num_envs=3)
    hp = torch.as_tensor(np.broadcast_to(g[f"{tag}_reset_hpos"], (3, Nh, 2)).copy(), device=env.device)
    hg = torch.as_tensor(np.broadcast_to(g[f"{tag}_reset_hgoal"], (3, Nh, 2)).copy(), device=env.device)
    S = env.reset(hp, hg)
    torch.cuda.synchronize()
    assert np.array_equal(S['HA']['detected_human_num'].cpu().numpy(), np.full(3, g[f"{tag}_reset_detected_humans"]))
    assert np.array_equal(S['HA']['human_valid_history'][0].cpu().numpy(), g[f"{tag}_reset_human_valid"])
    assert np.abs(S['HA']['spatial_edges'][0].cpu().numpy() - g[f"{tag}_reset_spatial_edges"]).max() <= 1e-9
