import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
E = 16384
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20), num_envs=E)
S = env.reset()
gen = torch.Generator(device=dev).manual_seed(1234)
noise = torch.randn((64, E, 2), generator=gen, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.2], device=dev, dtype=torch.float64)
act = S['PT']['MPC_actions'][:, :2].clone()
lo = torch.tensor([0.0, -1.0], device=dev, dtype=torch.float64); hi = torch.tensor([1.0, 1.0], device=dev, dtype=torch.float64)
agg = []
for t in range(80):
    torch.add(env.out["mpc_actions"][:, :2], noise[t % 64], out=act); torch.clamp(act, lo, hi, out=act)
    env.step(act)
    if t >= 40:
        agg.append(env.out["mpc_status"].clone())
st = torch.cat(agg).cpu().numpy()
for name, col in (("iterations", 0), ("retries", 1), ("lm_solves", 2), ("cost_evals", 3), ("term", 4)):
    v = st[:, col]
    qs = np.percentile(v, [50, 90, 99, 99.9, 100])
    print(name, "mean %.2f" % v.mean(), "p50/p90/p99/p99.9/max", qs)
ls = st[:, 2]
print("share of total lm_solves by problems with >20 solves: %.3f ; count frac %.4f" % (ls[ls > 20].sum() / ls.sum(), (ls > 20).mean()))
print("hist lm_solves", np.bincount(np.minimum(ls, 60))[:61])
