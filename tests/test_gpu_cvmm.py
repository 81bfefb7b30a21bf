"""GPU parity of the CVMM safety filter against the reference's own cvmm_safety_check / cvmm_diverse_safety_pipeline
(HA_and_PT env.py:422-513; tests/golden/cvmm.npz, produced by oracle/make_golden.py from the unmodified reference on
the shims). Every recorded query is one env of a batch. Exact: per-candidate safety flags and the selected action
(pure comparisons on fp64 roll-outs); the random fall-back (no safe action) is only checked to pick a back-up action."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cvmm.npz")
STATE = ("hpos", "hvel", "hgoal", "hstatic", "hhist", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
         "path_idx", "loc_to_start")


@pytest.mark.parametrize("tag,Nh", [("h6", 6), ("h10", 10)])
def test_cvmm_filter_matches_reference(tag, Nh):
    import torch
    import dr_mpc_b200 as d
    g = np.load(GOLDEN)
    Q = g[f"{tag}_action"].shape[0]
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=Q)
    env.reset()
    for name in STATE:
        a = np.ascontiguousarray(g[f"{tag}_pre_{name}"])
        env.state[name].copy_(torch.as_tensor(a, device=env.device).reshape(env.state[name].shape))
    before = {k: v.clone() for k, v in env.state.items()}
    acts = torch.as_tensor(g[f"{tag}_action"], device=env.device)
    safe, info = env.cvmm_diverse_safety_pipeline(acts, int(g["lookahead_steps"]), g["action_set"], want_candidates=True)
    torch.cuda.synchronize()
    cand = info['candidate_safe'].cpu().numpy()
    assert np.array_equal(cand, g[f"{tag}_cand_safe"]), np.argwhere(cand != g[f"{tag}_cand_safe"])[:5]
    rnd = g[f"{tag}_random"].astype(bool)
    assert np.array_equal(info['random'].cpu().numpy().astype(bool), rnd)
    out = safe.cpu().numpy()
    assert np.array_equal(out[~rnd], g[f"{tag}_safe_action"][~rnd])
    aset = g["action_set"]
    assert all((aset == row).all(1).any() for row in out[rnd])          # the fall-back is one of the back-up actions
    assert np.array_equal(info['original_safe'].cpu().numpy(), g[f"{tag}_cand_safe"][:, 0])
    kept = info['chosen'].cpu().numpy() == -1
    assert np.array_equal(kept, g[f"{tag}_cand_safe"][:, 0].astype(bool)) and np.array_equal(out[kept], g[f"{tag}_action"][kept])
    for k, v in env.state.items():                                       # a pure query
        assert torch.equal(v, before[k]), k
    print(tag, Q, "queries: original kept", int(kept.sum()), "filtered", int((~kept & ~rnd).sum()), "random", int(rnd.sum()))


def test_cvmm_filter_full_size_properties():
    """C3 size: a safe original action is returned untouched, a replaced one comes from the back-up set and is itself
    safe unless nothing is, and the answer is deterministic."""
    import torch
    import dr_mpc_b200 as d
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20), num_envs=16384)
    S = env.reset()
    for _ in range(25):
        S, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    a = S['PT']['MPC_actions'][:, :2].clone()
    s1, i1 = env.cvmm_diverse_safety_pipeline(a, want_candidates=True)
    s2, i2 = env.cvmm_diverse_safety_pipeline(a)
    assert torch.equal(s1, s2)
    kept = i1['chosen'] == -1
    assert torch.equal(s1[kept], a[kept]) and torch.equal(kept, i1['original_safe'] == 1)
    aset = torch.tensor(env.CVMM_ACTION_SET, dtype=torch.float64, device=env.device)
    rep = ~kept
    assert bool(rep.any())
    idx = i1['chosen'][rep].long()
    assert torch.equal(s1[rep], aset[idx])
    cand = i1['candidate_safe'][rep]
    picked_safe = cand.gather(1, (idx + 1).unsqueeze(1)).squeeze(1) == 1
    assert torch.equal(picked_safe, i1['random'][rep] == 0)
    # the pick has the highest speed among the safe back-up actions
    vs = aset[:, 0].unsqueeze(0).expand(cand.shape[0], -1).clone()
    vs[cand[:, 1:] == 0] = -1.0
    ok = i1['random'][rep] == 0
    assert torch.equal(aset[idx, 0][ok], vs.max(1).values[ok])
