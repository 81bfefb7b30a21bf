"""Timing probe (not a benchmark contract): per-kernel CUDA-event timings used to pick launch shapes."""
import json
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dr_mpc_b200 as d

dev = "cuda:0"
res = {}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def crowd(E, Nh, spread, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    ang = torch.rand((E, Nh), generator=g, device=dev, dtype=torch.float64) * 2 * np.pi
    rad = spread * torch.sqrt(torch.rand((E, Nh), generator=g, device=dev, dtype=torch.float64))
    hpos = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang)], -1)
    hvel = ((torch.rand((E, Nh, 2), generator=g, device=dev) - 0.5) * 1.2)
    rpos = (torch.rand((E, 2), generator=g, device=dev, dtype=torch.float64) - 0.5) * spread
    return hpos.contiguous(), hvel.contiguous(), (-hpos).contiguous(), rpos.contiguous()


for (E, Nh, spread) in ((1024, 10, 5.0), (16384, 20, 5.5), (65536, 50, 8.5), (8192, 50, 8.5), (16384, 6, 5.0)):
    hpos, hvel, goal, rpos = crowd(E, Nh, spread)
    for G in (4, 8, 16):
        o = d.BatchedORCA(d.EngineConfig(group_lanes=G))
        try:
            ms = timeit(lambda: o.predict(hpos, hvel, goal, rpos))
        except Exception as ex:
            ms = str(ex)
        res[f"orca_E{E}_N{Nh}_G{G}_ms"] = ms
        print(f"orca E={E} Nh={Nh} G={G}: {ms}", flush=True)

paths = d.paths.continuous_task_paths()
p0 = torch.tensor(paths[0], device=dev)
for (K, N) in ((4, 1024), (4, 16384), (4, 131072), (4, 1 << 20), (10, 131072), (20, 131072), (40, 131072)):
    g = torch.Generator(device=dev).manual_seed(K)
    node = torch.randint(0, 110, (N,), generator=g, device=dev)
    pose = p0[:, node].T.contiguous() + (torch.rand((N, 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.6
    interp = node.double()
    m = d.BatchedMPC(None, K, 0.25, True, 1.0, 1.0, num_problems=N)
    m.set_paths(paths)
    inp = {'interp_index': interp, 'pose': pose}
    # cold (iteration0) solve
    def cold():
        m.reset()
        m.act(inp)
    ms_cold = timeit(cold, iters=5, warm=1)
    m.act(inp, want_status=True)
    m.act(inp, want_status=True)
    st = m.last_status.cpu().numpy()
    ms_warm = timeit(lambda: m.act(inp), iters=5, warm=1)
    res[f"mpc_K{K}_N{N}"] = {"cold_ms": ms_cold, "warm_ms": ms_warm, "warm_iters": float(st[:, 0].mean()), "warm_lm_solves": float(st[:, 2].mean()),
                              "warm_cost_evals": float(st[:, 3].mean())}
    print(f"mpc K={K} N={N}: {res[f'mpc_K{K}_N{N}']}", flush=True)
    del m

for (E, Nh) in ((1024, 10), (16384, 20), (2048, 20), (65536, 20)):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
    S = env.reset()
    act = S['PT']['MPC_actions'][:, :2].clone()

    def step():
        act.copy_(env.out["mpc_actions"][:, :2])
        env.step(act)
    for _ in range(20):
        step()
    ms = timeit(step, iters=50, warm=5)
    c, s = env.stats()
    res[f"env_E{E}_N{Nh}_ms"] = ms
    print(f"env E={E} Nh={Nh}: {ms:.4f} ms/step -> {E / ms * 1e3:.3e} env-steps/s; stats {c.cpu().numpy()[:12]}", flush=True)
    if E == 16384:
        # split
        import ctypes
        from dr_mpc_b200 import _lib
        L = _lib.load()
        t0 = timeit(lambda: env.observe(False), iters=20)
        t1 = timeit(lambda: env.observe(True), iters=20)
        res["env_observe_ms"] = t0
        res["env_observe_mpc_ms"] = t1
        print("observe (ha obs + pt) ms", t0, "with mpc", t1, flush=True)
        # host path
        a_np = np.zeros((E, 2)); a_np[:, 0] = 0.5
        t = time.perf_counter()
        for _ in range(10):
            env.step_host(a_np)
        res["env_step_host_ms"] = (time.perf_counter() - t) / 10 * 1e3
        print("step_host ms", res["env_step_host_ms"], "slab MB", env.slab_bytes / 1e6, flush=True)
    del env

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
