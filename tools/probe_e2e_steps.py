"""step_host wall time per step against the MPC load of that step (where do the slow end-to-end steps come from?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
E, Nh = 16384, 20
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
S = env.reset()
rng = np.random.default_rng(0)
noise = rng.normal(size=(8, E, 2)) * [0.1, 0.2]
a_host = torch.empty((E, 2), dtype=torch.float64).pin_memory().numpy()
a_host[:] = S['PT']['MPC_actions'][:, :2].cpu().numpy()
rows = []
for i in range(260):
    t0 = time.perf_counter()
    v = env.step_host(a_host)
    t1 = time.perf_counter()
    np.clip(v["mpc_actions"][:, :2] + noise[i % 8], [0.0, -1.0], [1.0, 1.0], out=a_host)
    rows.append(((t1 - t0) * 1e3, float(v["mpc_status"][:, 2].mean()), float(v["mpc_status"][:, 3].mean())))
r = np.array(rows[60:])
print("mean %.3f p50 %.3f" % (r[:, 0].mean(), np.percentile(r[:, 0], 50)))
for lo, hi in ((0, 5), (5, 10), (10, 20), (20, 100)):
    m = (r[:, 1] >= lo) & (r[:, 1] < hi)
    if m.any(): print(f"lm_solves in [{lo},{hi}): {m.sum():3d} steps, step_host mean {r[m, 0].mean():.3f} ms (min {r[m, 0].min():.3f} max {r[m, 0].max():.3f}), cost_evals {r[m, 2].mean():.1f}")
