"""ORCA solve statistics of the C3 crowd over the first steps after reset (what the driver's bench window sees)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dr_mpc_b200 as d
E, Nh = 16384, 20
env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh), num_envs=E)
S = env.reset()
orca = d.BatchedORCA(env.cfg)
env.set_profiling(True)
for t in range(1, 61):
    env.kernel_ms(reset=True)
    S, *_ = env.step(S['PT']['MPC_actions'][:, :2].clone())
    kms, n = env.kernel_ms(reset=True)
    if t % 3 == 0 or t < 8:
        orca.predict(env.state["hpos"], env.state["hvel"], env.state["hgoal"], env.state["rpose"][:, :2].contiguous(), want_stats=True)
        torch.cuda.synchronize()
        st = orca.last_stats.double()
        lp3 = (st[..., 1] < st[..., 0]).double()
        print(f"step {t:3d} ha {kms['ha_step']:.3f} ms  mpc {kms['mpc_solve']:.3f}  lp3 frac {float(lp3.mean()):.3f}  n_lp1 {float(st[..., 2].mean()):.2f}  lp1_work {float(st[..., 3].mean()):.1f}  "
              f"proj/lp3-solve {float(st[..., 5].sum() / lp3.sum().clamp(min=1)):.2f}  fail line {float((st[..., 1] * lp3).sum() / lp3.sum().clamp(min=1)):.1f}  collide-branch frac {float(((st[..., 4].long() & 8) != 0).double().mean()):.3f}", flush=True)
