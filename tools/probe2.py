"""LP3 share + steady-state ORCA cost probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20), num_envs=16384)
S = env.reset()
act = S['PT']['MPC_actions'][:, :2].clone()
for t in range(120):
    act.copy_(env.out["mpc_actions"][:, :2]); env.step(act)
hp, hv, hg, rp = env.state["hpos"].clone(), env.state["hvel"].clone(), env.state["hgoal"].clone(), env.state["rpose"][:, :2].contiguous().clone()
for G in (8, 8 | 0x100, 8 | 0x200, 8 | 0x300):
    o = d.BatchedORCA(d.EngineConfig(group_lanes=G))
    ms = timeit(lambda: o.predict(hp, hv, hg, rp))
    o.predict(hp, hv, hg, rp, want_stats=True); torch.cuda.synchronize()
    st = o.last_stats
    print(f"steady-state crowd G={G&0xff} skip_lp3={bool(G&0x100)} static={bool(G&0x200)}: {ms:.4f} ms; LP3 frac {(st[...,1]<st[...,0]).double().mean().item():.3f}, mean n_lp1 {st[...,2].double().mean().item():.2f}, mean lp3 proj {st[...,5].double().mean().item():.2f}, max lp3 proj {st[...,5].max().item()}")
# fresh crowd right after reset (sparse ring)
env2 = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20), num_envs=16384); env2.reset()
hp, hv, hg, rp = env2.state["hpos"].clone(), env2.state["hvel"].clone(), env2.state["hgoal"].clone(), env2.state["rpose"][:, :2].contiguous().clone()
for G in (8, 8 | 0x100, 8 | 0x200, 8 | 0x300):
    o = d.BatchedORCA(d.EngineConfig(group_lanes=G))
    ms = timeit(lambda: o.predict(hp, hv, hg, rp))
    o.predict(hp, hv, hg, rp, want_stats=True); torch.cuda.synchronize()
    st = o.last_stats
    print(f"post-reset crowd G={G&0xff} skip_lp3={bool(G&0x100)} static={bool(G&0x200)}: {ms:.4f} ms; LP3 frac {(st[...,1]<st[...,0]).double().mean().item():.3f}, mean n_lp1 {st[...,2].double().mean().item():.2f}")
# pinned D2H bandwidth
x = torch.empty(59_000_000, dtype=torch.uint8, device=dev); h = torch.empty(59_000_000, dtype=torch.uint8).pin_memory()
ms = timeit(lambda: h.copy_(x, non_blocking=True), iters=10)
print("pinned D2H 59 MB:", ms, "ms ->", 59e6 / ms / 1e6, "GB/s")
