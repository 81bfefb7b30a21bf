"""End-to-end (host-buffer) step timing for different chunk counts of the step_host pipeline."""
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
for chunks in (1, 2, 4, 8, 16, 8):
    env.set_host_chunks(chunks)
    for i in range(5):
        v = env.step_host(a_host); np.clip(v["mpc_actions"][:, :2] + noise[i % 8], [0.0, -1.0], [1.0, 1.0], out=a_host)
    ts = []
    for i in range(60):
        t0 = time.perf_counter()
        v = env.step_host(a_host)
        t1 = time.perf_counter()
        np.clip(v["mpc_actions"][:, :2] + noise[i % 8], [0.0, -1.0], [1.0, 1.0], out=a_host)
        ts.append((t1 - t0) * 1e3)
    ts = np.array(ts)
    print(f"chunks {chunks:2d}: step_host mean {ts.mean():.3f} ms p10 {np.percentile(ts,10):.3f} p50 {np.percentile(ts,50):.3f} p90 {np.percentile(ts,90):.3f} "
          f"| lm_solves {float(v['mpc_status'][:,2].mean()):.1f}", flush=True)
# pinned D2H bandwidth
x = torch.empty(59_000_000, dtype=torch.uint8, device="cuda"); hbuf = torch.empty(59_000_000, dtype=torch.uint8).pin_memory()
for n in (59_000_000, 4_600_000, 500_000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): hbuf[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"pinned D2H {n/1e6:.1f} MB: {dt*1e3:.3f} ms -> {n/dt/1e9:.1f} GB/s")
