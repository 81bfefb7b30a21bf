"""Per-phase cycles of ha_step_kernel from a -DDRE_HA_TRACE build (DRE_LIB=variants/libdre_hatrace.so): thread 0 of every
CTA accumulates clock64 deltas into stat_counters[12..15]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dr_mpc_b200 as d
E, Nh = 16384, 20
for dedup in (False, True):
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, orca_dedup=dedup), num_envs=E)
    S = env.reset()
    for t in range(60):
        env.step(S['PT']['MPC_actions'][:, :2].clone())
    env.stats(reset=True)
    n = 40
    for t in range(n):
        env.step(S['PT']['MPC_actions'][:, :2].clone())
    c, _ = env.stats()
    c = c.cpu().numpy().astype(np.float64)
    ctas = (E + 5) // 6
    names = ("stage+pass1", "integrate+obs", "pass2", "rewards+merge+goals") if "trace2" not in os.environ.get("DRE_LIB", "") else ("per-human terms", "per-env merge+stats", "goal detect", "goal resample")
    tot = c[12:16].sum()
    print("dedup" if dedup else "two-pass", "| cycles per CTA per step:", {nm: int(v / n / ctas) for nm, v in zip(names, c[12:16])},
          "| shares:", {nm: round(v / tot, 3) for nm, v in zip(names, c[12:16])})
