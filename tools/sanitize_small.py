"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
for cfg in (dict(num_humans=6), dict(num_humans=20, soft_reset_on_device=True, orca_dedup=True, static_humans=((-4.0, 0.0),))):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(**cfg), num_envs=96 + 5)
    S = env.reset()
    for t in range(4):
        a = S['PT']['MPC_actions'][:, :2].clone()
        a2, _ = env.cvmm_diverse_safety_pipeline(a)
        S, *_ = env.step(a2)
    env.set_host_chunks(3)
    v = env.step_host(a.cpu().numpy())
    env.observe(True)
    torch.cuda.synchronize()
o = d.BatchedORCA(d.EngineConfig())
g = torch.Generator(device="cuda").manual_seed(0)
hp = (torch.rand((33, 50, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 10
o.predict(hp, torch.zeros((33, 50, 2), device="cuda"), -hp, torch.zeros((33, 2), device="cuda", dtype=torch.float64), want_stats=True)
paths = d.paths.continuous_task_paths()
for K in (4, 10, 40):
    m = d.BatchedMPC(None, K, 0.25, True, 1.0, 1.0, num_problems=70); m.set_paths(paths)
    p0 = torch.tensor(paths[0], device="cuda")
    node = torch.randint(0, 100, (70,), device="cuda")
    m.act({'interp_index': node.double(), 'pose': p0[:, node].T.contiguous() + 0.1}, want_status=True)
torch.cuda.synchronize()
print("sanitize run done")
