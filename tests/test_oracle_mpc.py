"""CPU: the flat planar LM oracle (oracle/mpc_ref.c, block-tridiagonal SE(2)) against golden sequences produced by the
reference's own MPC.act + LevMarqGaussNewtonCustomSolver (dense SE(3)) running on the pysteam/pylgmath shims."""
import numpy as np
import pytest
from conftest import mpc_path_table

# controls agree to TOL when both solvers take the same LM decision sequence; near the end of a path (v pinned at the
# lower bound by the feasibility scaling) the reference itself is chaotic (1e-13 input perturbations move its output by
# 1e-3), so those steps are counted as branch divergences and only bounded loosely.
TOL = {4: 1e-6, 10: 2e-6}
LOOSE = 0.2


@pytest.mark.parametrize("K", [4, 10])
def test_oracle_matches_reference_teacher_forced(oracle, golden_mpc, K):
    g = golden_mpc
    paths, tab, lens = mpc_path_table(g)
    n = g[f"K{K}_u"].shape[0]
    mb = oracle.MpcBatch(n, oracle.mpc_params(K=K), tab, lens)
    mb.warm_T[:] = g[f"K{K}_warm_T"]; mb.warm_u[:] = g[f"K{K}_warm_u"]; mb.iteration0[:] = g[f"K{K}_it0"]
    u, st = mb.act(g[f"K{K}_path_id"], g[f"K{K}_interp"], g[f"K{K}_pose"])
    d = np.abs(u - g[f"K{K}_u"]).max(axis=1)
    same = st["iterations"] == g[f"K{K}_iters"]
    regular = g[f"K{K}_interp"] < lens[g[f"K{K}_path_id"]] - 1 - 1.2375 * K      # no clamped reference yet
    tight = d <= TOL[K]
    print(f"K={K}: {n} solves, max|du| same-branch={d[same & regular].max():.2e}, divergences={int((~tight).sum())}, "
          f"max|du| overall={d.max():.2e}")
    assert tight[regular & same].all()
    assert (~tight).sum() <= 0.12 * n and d.max() < LOOSE
    assert (st["term"] > 0).all()
    # final costs agree wherever the controls do
    assert np.allclose(st["final_cost"][tight], g[f"K{K}_cost"][tight], rtol=1e-6, atol=1e-9)


def test_iteration0_initial_guess_and_warm_start_shift(oracle, golden_mpc):
    g = golden_mpc
    paths, tab, lens = mpc_path_table(g)
    mb = oracle.MpcBatch(1, oracle.mpc_params(K=4), tab, lens)
    u0, _ = mb.act([0], [3.2], [paths[0][:, 3] + [0.05, 0.02, 0.01]])
    assert mb.iteration0[0] == 0
    # controls shifted, last control guess (0.5, 0); poses NOT shifted: last two guesses are equal (mpc.py:196-205)
    assert np.allclose(mb.warm_u[0, :3].ravel(), u0[0, 2:]) and np.allclose(mb.warm_u[0, 3], [0.5, 0.0])
    assert np.array_equal(mb.warm_T[0, 2], mb.warm_T[0, 3])


def test_controls_respect_bounds(oracle, golden_mpc):
    g = golden_mpc
    paths, tab, lens = mpc_path_table(g)
    rng = np.random.default_rng(5)
    N = 512
    pid = rng.integers(0, 4, N).astype(np.int32)
    idx = rng.uniform(0, lens[pid] - 1)
    lo = np.floor(idx).astype(int)
    pose = np.stack([tab[pid, r, lo] for r in range(3)], 1) + rng.uniform(-0.4, 0.4, (N, 3))
    mb = oracle.MpcBatch(N, oracle.mpc_params(K=4), tab, lens)
    u, st = mb.act(pid, idx, pose)
    v, w = u[:, 0::2], u[:, 1::2]
    assert (v > 0).all() and (v < 1).all() and (np.abs(w) < 1).all()
