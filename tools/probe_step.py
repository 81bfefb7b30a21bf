"""Per-kernel timing of the C3 env step under the bench's closed loop (not a benchmark contract). Variants come from env
vars read by libdre.so (DRE_MPC_MINB, ...). usage: python tools/probe_step.py [tag] [steps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
tag = sys.argv[1] if len(sys.argv) > 1 else "base"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
E, Nh = int(os.environ.get("PROBE_E", 16384)), int(os.environ.get("PROBE_NH", 20))
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, group_lanes=int(os.environ.get("PROBE_G", 0)), orca_dedup=bool(int(os.environ.get("PROBE_DEDUP", 0)))), num_envs=E)
S = env.reset()
gen = torch.Generator(device=dev).manual_seed(1234)
noise = torch.randn((64, E, 2), generator=gen, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.2], device=dev, dtype=torch.float64)
act = S['PT']['MPC_actions'][:, :2].clone()
lo = torch.tensor([0.0, -1.0], device=dev, dtype=torch.float64); hi = torch.tensor([1.0, 1.0], device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
env.set_profiling(True)
out = {"tag": tag, "windows": []}
t = 0
chk = 0.0
for w in range(steps // 50):
    env.kernel_ms(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(50):
        torch.add(env.out["mpc_actions"][:, :2], noise[t % 64], out=act); torch.clamp(act, lo, hi, out=act)
        t += 1
        env.step(act); flush.zero_()
    kms, n = env.kernel_ms(reset=True)
    st = env.out["mpc_status"].double()
    row = {k: round(v / n, 4) for k, v in kms.items()}
    row["lm_solves"] = round(float(st[:, 2].mean()), 2); row["retries"] = int(st[:, 1].sum())
    out["windows"].append(row)
    print(tag, "steps", t, row, flush=True)
chk = float(env.out["rewards"].sum()) + float(env.state["hpos"].sum())
cnt, sums = env.stats()
out["checksum"] = chk; out["stats"] = cnt.cpu().numpy().tolist()
print(tag, "checksum %.12e" % chk, "stats", cnt.cpu().numpy()[:12].tolist(), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/probe_step_{tag}.json", "w"))
