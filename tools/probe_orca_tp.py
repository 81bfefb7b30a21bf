"""Thread-per-agent ORCA pass (group_lanes = 1, +0x400: projected lines in global scratch) against the 8-lane kernel:
bit-equality of the velocities and pass time, on a random crowd and on states the env reaches."""
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
E_, Nh = int(os.environ.get("PROBE_E", 16384)), int(os.environ.get("PROBE_NH", 20))
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E_)
S = env.reset()
for t in range(40):
    S, R, D, info, eps = env.step(S['PT']['MPC_actions'][:, :2].clone())
torch.cuda.synchronize()
hpos = env.state["hpos"].clone(); hvel = env.state["hvel"].clone(); goal = env.state["hgoal"].clone(); rpos = env.state["rpose"][:, :2].contiguous().clone()
ref = None
for G in (8, 1, 1 | 0x400):
    o = d.BatchedORCA(d.EngineConfig(group_lanes=G))
    v = o.predict(hpos, hvel, goal, rpos).clone()
    ms = timeit(lambda: o.predict(hpos, hvel, goal, rpos))
    if ref is None: ref = v
    same = torch.equal(v.view(torch.int32), ref.view(torch.int32))
    print(f"G={G:#x}: {ms:.4f} ms per pass ({E_}x{Nh}), bit-identical to G=8: {same}, mismatching agents {(v != ref).any(-1).sum().item()}", flush=True)
o = d.BatchedORCA(d.EngineConfig(group_lanes=8))
o.predict(hpos, hvel, goal, rpos, want_stats=True)
st = o.last_stats.cpu().numpy().reshape(-1, 6).astype(np.int64)
lp3 = st[:, 1] < st[:, 0]
print(f"agents {len(st)}: lines mean {st[:,0].mean():.1f}, LP1 calls mean {st[:,2].mean():.2f} (max {st[:,2].max()}), LP1 scan work mean {st[:,3].mean():.1f}, "
      f"LP3 agents {lp3.mean()*100:.1f} %, projections per LP3 agent {st[lp3,5].mean():.2f} (max {st[:,5].max()}), fail line mean {st[lp3,1].mean():.1f}")
