"""CPU: the flat planar LM oracle (oracle/mpc_ref.c, block-tridiagonal SE(2)) against golden sequences produced by the
reference's own MPC.act + LevMarqGaussNewtonCustomSolver (dense SE(3)) running on the pysteam/pylgmath shims."""
import numpy as np
import pytest
from conftest import mpc_path_table

# Measured envelope (this container, gcc -O2 -ffp-contract=off): max |du| 4.3e-8 (K=4) / 5.0e-8 (K=10) / 9e-8 (K=20, 40) over
# every golden solve, the outer-iteration count equal on every one. The assert is that envelope with one decade of head
# room: EVERY solve within TOL and ZERO LM-iteration divergences, except the solves listed in ALLOWED_DIVERGENT (golden
# index per horizon; empty today). Near the end of a path (v pinned at its lower bound by the feasibility scaling) the
# reference itself is chaotic (a 1e-13 input perturbation moves its output by 1e-3), which is why the list exists at all.
TOL = {4: 1e-6, 10: 1e-6, 20: 1e-6, 40: 1e-6}
ALLOWED_DIVERGENT = {4: (), 10: (), 20: (), 40: ()}


@pytest.mark.parametrize("K", [4, 10, 20, 40])
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
    ok = np.ones(n, dtype=bool)
    ok[list(ALLOWED_DIVERGENT[K])] = False
    assert tight[ok].all() and same[ok].all(), (np.nonzero(~tight | ~same)[0], d.max())
    assert (~regular).any()                      # the end-of-path clamp of the reference poses is in the fixture
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


@pytest.mark.parametrize("K", [4, 10])
def test_oracle_forced_failures_take_the_reference_retry_path(oracle, golden_mpc, K):
    """Solves whose LM step fails, recorded from the reference's own MPC.act: its `except: pose_error_scaling *= 5; continue`
    (mpc.py:188-190) ran on every one. kind 0: a failing FIRST solve after reset() -- the retry starts from the warm start
    stored BEFORE the reset (mpc.py:72-75 clears iteration0 before the solve); kind 1: a failing warm-started solve."""
    g = golden_mpc
    paths, tab, lens = mpc_path_table(g)
    n = g[f"F{K}_u"].shape[0]
    mb = oracle.MpcBatch(n, oracle.mpc_params(K=K), tab, lens)
    mb.warm_T[:] = g[f"F{K}_warm_T"]; mb.warm_u[:] = g[f"F{K}_warm_u"]; mb.iteration0[:] = g[f"F{K}_it0"]
    u, st = mb.act(g[f"F{K}_path_id"], g[f"F{K}_interp"], g[f"F{K}_pose"])
    d = np.abs(u - g[f"F{K}_u"]).max(axis=1)
    print(f"K={K}: {n} failing solves ({int((g[f'F{K}_kind'] == 0).sum())} after reset), max|du|={d.max():.2e}, retries {st['retries'].tolist()}")
    assert (g[f"F{K}_retries"] >= 1).all() and (g[f"F{K}_kind"] == 0).any() and (g[f"F{K}_kind"] == 1).any()
    assert np.array_equal(st["retries"], g[f"F{K}_retries"])
    assert np.array_equal(st["iterations"], g[f"F{K}_iters"])
    assert d.max() <= 1e-6
    assert np.allclose(st["final_cost"], g[f"F{K}_cost"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("K", [4, 10, 20])
def test_closed_loop_tracks_every_path(oracle, K):
    """A property no restatement can fake: the controls, fed to the exact-arc unicycle (unicycle.py:89-126) in closed loop
    with a fresh localisation every step, pull a robot that starts 18 cm / 0.2 rad off onto each of the four default paths,
    keep it within 6 cm (measured: 1.3 - 3.8 cm) at 0.7 - 0.95 m/s and bring it to the last node, inside the action box."""
    import dr_mpc_b200.paths as P
    paths = P.continuous_task_paths()
    tab, lens = P.pack_paths(paths)

    def localize(path, pose):                  # nearest node + projection on the adjacent segment (path_tracking_core.py:21-81)
        d = np.hypot(path[0] - pose[0], path[1] - pose[1])
        i = int(np.argmin(d))
        n = path.shape[1]
        j = i + 1 if (i == 0 or (i + 1 < n and d[i + 1] < d[i - 1])) else i - 1
        a, b = min(i, j), min(max(i, j), n - 1)
        if a == b:
            return float(a), d[i]
        seg = path[:2, b] - path[:2, a]
        t = np.clip(np.dot(pose[:2] - path[:2, a], seg) / np.dot(seg, seg), 0.0, 1.0)
        return a + t, float(np.hypot(*(pose[:2] - (path[:2, a] + t * seg))))

    def unicycle(pose, v, w, dt=0.25):
        x, y, th = pose
        dx, dy = (v * dt, 0.0) if w == 0 else (v / w * np.sin(w * dt), v / w * (1 - np.cos(w * dt)))
        return np.array([x + np.cos(th) * dx - np.sin(th) * dy, y + np.sin(th) * dx + np.cos(th) * dy,
                         (th + w * dt + np.pi) % (2 * np.pi) - np.pi])

    for pid, path in enumerate(paths):
        m = oracle.MpcBatch(1, oracle.mpc_params(K=K), tab, lens)
        pose = path[:, 0] + np.array([0.15, -0.1, 0.2])
        devs, speeds, idx = [], [], 0.0
        for _ in range(160):
            idx, dev = localize(path, pose)
            if idx >= path.shape[1] - 1 - 1e-9:
                break
            u, st = m.act(np.array([pid], np.int32), np.array([idx]), pose[None])
            assert st["term"][0] >= 0 and np.isfinite(u).all()
            assert (u[0, 0::2] >= 0).all() and (u[0, 0::2] <= 1.0).all() and (np.abs(u[0, 1::2]) <= 1.0).all()
            pose = unicycle(pose, u[0, 0], u[0, 1])
            devs.append(dev)
            speeds.append(u[0, 0])
        assert idx >= path.shape[1] - 1 - 1e-9, (K, pid, idx)                  # arrived at the last node
        assert max(devs[10:]) < 0.06, (K, pid, max(devs[10:]))
        assert np.mean(speeds[5:]) > 0.65, (K, pid, np.mean(speeds[5:]))
