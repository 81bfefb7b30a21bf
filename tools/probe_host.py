"""tools/probe_host.py: the host-buffer step (dre_env_step_host) in the driver's window with the library's own host-side
timeline (DRE_HOST_TRACE=1 prints one call in 64 to stderr). usage: DRE_HOST_TRACE=1 python tools/probe_host.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dr_mpc_b200 as d
E, Nh = int(os.environ.get("PROBE_E", 16384)), int(os.environ.get("PROBE_NH", 20))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for rep in range(3):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
    env.reset()
    if os.environ.get("PROBE_CHUNKS"):
        d._lib.check(env._L.dre_env_set_host_chunks(env._h, int(os.environ["PROBE_CHUNKS"])))
    a = torch.empty((E, 2), dtype=torch.float64).pin_memory().numpy()
    a[:] = env.out["mpc_actions"][:, :2].cpu().numpy()
    rng = np.random.default_rng(0)
    noise = rng.normal(size=(8, 2, E)) * np.array([0.1, 0.2])[None, :, None]

    def policy(v, i):
        ma = v["mpc_actions"]
        np.add(ma[:, 0], noise[i % 8, 0], out=a[:, 0]); np.clip(a[:, 0], 0.0, 1.0, out=a[:, 0])
        np.add(ma[:, 1], noise[i % 8, 1], out=a[:, 1]); np.clip(a[:, 1], -1.0, 1.0, out=a[:, 1])

    for i in range(5):
        policy(env.step_host(a), i)
    t0 = time.perf_counter()
    tp = 0.0
    for i in range(steps):
        v = env.step_host(a)
        t1 = time.perf_counter()
        policy(v, i)
        tp += time.perf_counter() - t1
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt / steps * 1e3:.4f} ms/step ({E * steps / dt * 1e-6:.2f} M env-steps/s), host policy {tp / steps * 1e6:.0f} us/step", flush=True)
    del env
