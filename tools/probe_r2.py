"""Round-2 A/B probe (not a benchmark contract): (1) the driver's window of the C3 step (5 warm-up + 20 timed steps right after
reset), (2) per-kernel event times in 25-step windows, (3) a state checksum, (4) the C4 ORCA-only step at neighbor_dist
10 and 3. usage: python tools/probe_r2.py [tag] [windows]   (DRE_LIB selects another build of libdre.so)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dr_mpc_b200 as d
dev = "cuda:0"
tag = sys.argv[1] if len(sys.argv) > 1 else "base"
windows = int(sys.argv[2]) if len(sys.argv) > 2 else 6
E, Nh = int(os.environ.get("PROBE_E", 16384)), int(os.environ.get("PROBE_NH", 20))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lo = torch.tensor([0.0, -1.0], device=dev, dtype=torch.float64); hi = torch.tensor([1.0, 1.0], device=dev, dtype=torch.float64)
gen = torch.Generator(device=dev).manual_seed(1234)
noise = torch.randn((64, E, 2), generator=gen, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.2], device=dev, dtype=torch.float64)
out = {"tag": tag}


def make():
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, group_lanes=int(os.environ.get('PROBE_G', 0)), orca_lookahead_prune=bool(int(os.environ.get('PROBE_PRUNE', 0))), orca_dedup=bool(int(os.environ.get('PROBE_DEDUP', 0)))), num_envs=E)
    env.reset()
    return env, env.out["mpc_actions"][:, :2].clone(), [0]


def step(env, act, t):
    torch.add(env.out["mpc_actions"][:, :2], noise[t[0] % 64], out=act); torch.clamp(act, lo, hi, out=act)
    t[0] += 1
    env.step(act)


# (1) the driver's window
vals = []
for rep in range(3):
    env, act, t = make()
    for _ in range(5):
        step(env, act, t); flush.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(); step(env, act, t); b.record(); flush.zero_()
    torch.cuda.synchronize()
    vals.append(sum(a.elapsed_time(b) for a, b in ev) / 20)
    del env
out["driver_window_ms"] = vals
print(tag, "driver window (5+20 steps after reset) ms/step:", ["%.4f" % v for v in vals], "-> %.2f M env-steps/s" % (E / min(vals) * 1e-3), flush=True)

if os.environ.get('PROBE_ONLY_WINDOW'):
    sys.exit(0)
# (2) per-kernel windows
env, act, t = make()
env.set_profiling(True)
out["windows"] = []
for w in range(windows):
    env.kernel_ms(reset=True)
    for _ in range(25):
        step(env, act, t); flush.zero_()
    kms, n = env.kernel_ms(reset=True)
    st = env.out["mpc_status"].double()
    row = {k: round(v / n, 4) for k, v in kms.items()}
    row["lm_solves"] = round(float(st[:, 2].mean()), 2); row["retries"] = int(st[:, 1].sum())
    out["windows"].append(row)
    print(tag, "steps", t[0], row, flush=True)
env.set_profiling(False)
# unprofiled steady window (kernels overlap)
for _ in range(10):
    step(env, act, t); flush.zero_()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
torch.cuda.synchronize()
for a, b in ev:
    a.record(); step(env, act, t); b.record(); flush.zero_()
torch.cuda.synchronize()
out["steady_ms"] = sum(a.elapsed_time(b) for a, b in ev) / 50
chk = float(env.out["rewards"].sum()) + float(env.state["hpos"].sum())
cnt, sums = env.stats()
out["checksum"] = chk
print(tag, "steady (steps %d..%d) %.4f ms/step; checksum %.12e stats %s" % (t[0] - 50, t[0], out["steady_ms"], chk, cnt.cpu().numpy()[:12].tolist()), flush=True)
del env

# (4) C4 shard: ORCA-only step, 8192 x 50
for nd in (10.0, 3.0):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=50, neighbor_dist=nd, circle_radius_humans=8.0), num_envs=8192)
    env.reset()
    for _ in range(5):
        env.crowd_step(); flush.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(); env.crowd_step(); b.record(); flush.zero_()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / 20
    out[f"c4_nd{nd:g}_ms"] = ms
    print(tag, f"C4 8192x50 neighbor_dist {nd:g}: {ms:.4f} ms/step, hpos checksum {float(env.state['hpos'].sum()):.10e}", flush=True)
    del env
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/probe_r2_{tag}.json", "w"))
