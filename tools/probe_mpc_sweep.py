"""C5: batched MPC solve sweep over the horizon (1 cold solve + 4 warm re-solves after applying the first control)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
N = int(os.environ.get("PROBE_N", 1 << 20))
paths = d.paths.continuous_task_paths()
p0 = torch.tensor(paths[0], device=dev)
res = {}
# process warm-up (context, module load, allocator): a small throw-away batch
_w = d.BatchedMPC(None, 4, 0.25, True, 1.0, 1.0, num_problems=4096); _w.set_paths(paths)
for _ in range(3): _w.act({'interp_index': torch.zeros(4096, device=dev, dtype=torch.float64), 'pose': p0[:, :1].T.repeat(4096, 1).contiguous()})
torch.cuda.synchronize(); del _w
for K in [int(x) for x in os.environ.get("PROBE_KS", "4,10,20,30,40").split(",")]:
    n = N
    g = torch.Generator(device=dev).manual_seed(K)
    node = torch.randint(0, 110, (n,), generator=g, device=dev)
    pose = p0[:, node].T.contiguous() + (torch.rand((n, 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.6
    interp = node.double()
    m = d.BatchedMPC(None, K, 0.25, True, 1.0, 1.0, num_problems=n)
    m.set_paths(paths)
    ts, its = [], []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        u = m.act({'interp_index': interp, 'pose': pose}, want_status=True)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)); its.append(float(m.last_status[:, 2].double().mean()))
        # apply the first control (exact-arc unicycle), advance the localisation by the travelled nodes
        v, w = u[:, 0], u[:, 1]
        th = pose[:, 2]
        dx = torch.where(w.abs() > 0, v / w * torch.sin(w * 0.25), v * 0.25); dy = torch.where(w.abs() > 0, v / w * (1 - torch.cos(w * 0.25)), torch.zeros_like(v))
        pose = torch.stack([pose[:, 0] + torch.cos(th) * dx - torch.sin(th) * dy, pose[:, 1] + torch.sin(th) * dx + torch.cos(th) * dy, th + w * 0.25], 1).contiguous()
        interp = interp + v * 0.25 / 0.2
    res[K] = {"problems": n, "cold_ms": ts[0], "warm_ms": float(np.mean(ts[2:])), "cold_lm_solves": its[0], "warm_lm_solves": float(np.mean(its[2:])),
              "warm_solves_per_s": n / (np.mean(ts[2:]) * 1e-3)}
    print(K, res[K], flush=True)
    del m
json.dump(res, open("gpurun_out/probe_mpc_sweep.json", "w"), indent=1)
