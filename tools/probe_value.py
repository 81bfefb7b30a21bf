"""Device-resident step time (bench's `value` loop) for A/B runs of the step schedule."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
E, Nh = 16384, 20
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
S = env.reset()
gen = torch.Generator(device=dev).manual_seed(1234)
noise = torch.randn((64, E, 2), generator=gen, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.2], device=dev, dtype=torch.float64)
act = S['PT']['MPC_actions'][:, :2].clone()
lo = torch.tensor([0.0, -1.0], device=dev, dtype=torch.float64); hi = torch.tensor([1.0, 1.0], device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
t = 0
for w in range(6):
    tot = 0.0
    for _ in range(50):
        torch.add(env.out["mpc_actions"][:, :2], noise[t % 64], out=act); torch.clamp(act, lo, hi, out=act); t += 1
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(act); b.record(); flush.zero_()
        torch.cuda.synchronize(); tot += a.elapsed_time(b)
    print(sys.argv[1] if len(sys.argv) > 1 else "", "steps", t, "ms/step %.4f" % (tot / 50), "lm_solves %.1f" % float(env.out["mpc_status"][:, 2].double().mean()), flush=True)
print("checksum %.10e" % (float(env.out["rewards"].sum()) + float(env.state["hpos"].sum())))
