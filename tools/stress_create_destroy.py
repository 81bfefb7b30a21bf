import sys; sys.path.insert(0, '/root/repo')
import torch, dr_mpc_b200 as d, gc
torch.zeros(1, device='cuda')
free0 = torch.cuda.mem_get_info()[0]
for i in range(40):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=20, soft_reset_on_device=bool(i % 2), orca_dedup=bool(i % 3 == 0)), num_envs=4096)
    S = env.reset()
    for t in range(3):
        S, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    env.step_host(S['PT']['MPC_actions'][:, :2].cpu().numpy())
    del env, S
    gc.collect()
torch.cuda.synchronize()
free1 = torch.cuda.mem_get_info()[0]
print("device memory drift after 40 create/destroy cycles: %.1f MB" % ((free0 - free1) / 1e6))
