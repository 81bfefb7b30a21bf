"""Tiny driver for ncu captures: C3 shape, reset + N free-running steps (argv[1], default 12)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dr_mpc_b200 as d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20), num_envs=16384)
S = env.reset()
for t in range(n):
    S, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
torch.cuda.synchronize()
print("ok", float(env.out["rewards"].sum()))
