import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import dr_mpc_b200 as d
E, Nh = 16384, 20
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
S = env.reset()
rng = np.random.default_rng(0)
noise = rng.normal(size=(8, E, 2)) * [0.1, 0.2]
a_host = torch.empty((E, 2), dtype=torch.float64).pin_memory().numpy()
a_host[:] = S['PT']['MPC_actions'][:, :2].cpu().numpy()
lo_h, hi_h = np.array([0.0, -1.0]), np.array([1.0, 1.0])
ts, tp = [], []
for i in range(260):
    t0 = time.perf_counter()
    v = env.step_host(a_host)
    t1 = time.perf_counter()
    np.add(v["mpc_actions"][:, :2], noise[i % 8], out=a_host); np.clip(a_host, lo_h, hi_h, out=a_host)
    t2 = time.perf_counter()
    ts.append(t1 - t0); tp.append(t2 - t1)
ts, tp = np.array(ts[60:]) * 1e3, np.array(tp[60:]) * 1e3
print("step_host mean %.3f ms, policy mean %.3f ms (p50 %.3f)" % (ts.mean(), tp.mean(), np.percentile(tp, 50)))
