"""Distribution of LM work per problem over a lap (where does the MPC kernel's time go in the end-of-path storms?)."""
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
env.set_profiling(True)
for t in range(160):
    torch.add(env.out["mpc_actions"][:, :2], noise[t % 64], out=act); torch.clamp(act, lo, hi, out=act)
    env.kernel_ms(reset=True)
    env.step(act)
    kms, n = env.kernel_ms(reset=True)
    st = env.out["mpc_status"].cpu().numpy()   # int32 view [E, 8]: iterations, retries, lm_solves, cost_evals, term, zero_steps, cost(lo,hi)
    ls = st[:, 2].astype(np.int64); it = st[:, 0]; rt = st[:, 1]; zs = st[:, 5]; ce = st[:, 3]
    if t >= 95:
        print(f"t={t:3d} mpc {kms['mpc_solve']:.3f} ms | lm_solves mean {ls.mean():6.2f} p50 {np.percentile(ls,50):.0f} p99 {np.percentile(ls,99):.0f} max {ls.max()} | "
              f"iters mean {it.mean():.2f} max {it.max()} | retries sum {rt.sum()} max {rt.max()} | zero_steps sum {zs.sum()} | cost_evals mean {ce.mean():.1f} | "
              f"idx mean {float(env.out['info'][:,5].mean()):.1f} path {int(env.state['path_idx'][0])} | executed attempts est {(ce/2).mean():.1f}", flush=True)
