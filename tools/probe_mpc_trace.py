"""DRE_MPC_TRACE=1 python tools/probe_mpc_trace.py : per-phase cycle trace of the MPC kernel at chosen steps of the bench loop."""
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
for t in range(152):
    torch.add(env.out["mpc_actions"][:, :2], noise[t % 64], out=act); torch.clamp(act, lo, hi, out=act)
    if t in (60, 61, 135, 136, 148, 149, 150):
        sys.stderr.write(f"t={t} "); sys.stderr.flush()
    else:
        sys.stderr.write("skip "); sys.stderr.flush()
    env.step(act)
    torch.cuda.synchronize()
