"""Stand-alone ORCA pass timing (C2 / C3 / C4 shapes) for launch-shape choices."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
dev = "cuda:0"
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
def crowd(E, Nh, spread, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    ang = torch.rand((E, Nh), generator=g, device=dev, dtype=torch.float64) * 2 * np.pi
    rad = spread * torch.sqrt(torch.rand((E, Nh), generator=g, device=dev, dtype=torch.float64))
    hpos = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang)], -1)
    hvel = ((torch.rand((E, Nh, 2), generator=g, device=dev) - 0.5) * 1.2)
    rpos = (torch.rand((E, 2), generator=g, device=dev, dtype=torch.float64) - 0.5) * spread
    return hpos.contiguous(), hvel.contiguous(), (-hpos).contiguous(), rpos.contiguous()
for (E, Nh, spread) in ((1024, 10, 5.0), (16384, 20, 5.5), (8192, 50, 8.5), (65536, 50, 8.5)):
    hpos, hvel, goal, rpos = crowd(E, Nh, spread)
    for G in (4, 8, 16):
        o = d.BatchedORCA(d.EngineConfig(group_lanes=G))
        ms = timeit(lambda: o.predict(hpos, hvel, goal, rpos), iters=10)
        print(f"orca E={E} Nh={Nh} G={G}: {ms:.4f} ms -> {E * Nh / ms * 1e-6:.1f} G solves/s... {E / ms * 1e3:.3e} env-passes/s", flush=True)
