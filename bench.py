#!/usr/bin/env python
"""bench.py -- throughput of the DR-MPC hot path (ORCA crowd + MPC solve + env step) on N B200s.

    python bench.py --gpus N --steps K --warmup W                      # our arm, BASELINE configs[2] (C3), weak scaling
    python bench.py --config {C2,C3,C4,C5} --scaling {weak,strong} ... # the other BASELINE configs (profiles/configs_r2.jsonl)
    python bench.py --impl reference --gpus N --steps K --warmup W     # the CPU path on the box's host cores

Configs (BASELINE.json.configs / SURVEY 8(d)):
  C2  1024 envs x 10 humans, K = 4, full env step (ORCA + MPC horizon step), one GPU
  C3  16384 envs x 20 humans, K = 4, full env step. weak: 16384 envs PER GPU (default, the driver's line);
      strong: 16384 envs in total, E/G per GPU
  C4  65536 envs x 50-human dense crowd, ORCA-only step (one pass + integration), sharded over the GPUs (strong: 65536/G
      per GPU; weak: 8192 per GPU = the 8-GPU shard); --neighbor-dist 3 exercises the neighbour cut-off
  C5  MPC solve sweep: 1 M problems, horizon --horizon in 10..40, iteration0 solve + 4 warm re-solves (strong: 1M/G per GPU)
One "step" = one pass of the config's hot path over every env / problem of the shard. Before timing, every config runs a
PARITY GATE: a sample of its own device state at that exact shape goes through the CPU oracle and the mismatch / LP-branch
/ LM-branch divergence counts are printed in the line (`parity`).

Prints ONE JSON line (rank 0). `value` = device-timed whole-job throughput with state resident in HBM; `e2e` = the same
through the host-buffer C-ABI call (H2D of the inputs + D2H of the whole result every step); `roofline` /
`roofline_compute` / `kernels_ms` explain it; `cpu_baseline` = the CPU oracle port timed on this box's host cores on a
bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRICS = {"C2": ("env-steps/sec (ORCA crowd + MPC solve)", "env-steps/s"), "C3": ("env-steps/sec (ORCA crowd + MPC solve)", "env-steps/s"),
           "C4": ("env-steps/sec (ORCA-only crowd step)", "env-steps/s"), "C5": ("MPC solves/sec (LM, warm re-solves)", "solves/s")}
DTYPE = "f32 ORCA + f64 MPC/env"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="override the config's shard size")
    ap.add_argument("--humans", type=int, default=0)
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--neighbor-dist", type=float, default=10.0)
    ap.add_argument("--cpu-envs", type=int, default=0, help="envs / problems of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dedup", action="store_true", help="skip the separately reported orca_dedup variant")
    ap.add_argument("--no-parity-gate", action="store_true")
    ap.add_argument("--soft-reset", action="store_true", help="C2/C3: run HAAndPTEnv.soft_reset's recovery machine on device")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    total = {"C2": 1024, "C3": 16384, "C4": 65536, "C5": 1 << 20}[a.config]
    a.humans = a.humans or {"C2": 10, "C3": 20, "C4": 50, "C5": 0}[a.config]
    a.horizon = a.horizon or {"C2": 4, "C3": 4, "C4": 4, "C5": 10}[a.config]
    if not a.envs_per_gpu:
        if a.scaling == "strong":
            a.envs_per_gpu = (total + world - 1) // world
        else:
            a.envs_per_gpu = {"C2": 1024, "C3": 16384, "C4": 8192, "C5": 1 << 17}[a.config]
    a.cpu_envs = a.cpu_envs or {"C2": 1024, "C3": 4096, "C4": 1024, "C5": 16384}[a.config]
    return a


def workload_config(a):
    E, Nh, K = a.envs_per_gpu, a.humans, a.horizon
    what = {"C2": f"C2 full env step: {E} envs x {Nh} humans per GPU, MPC horizon {K}",
            "C3": f"C3 full env step: {E} envs x {Nh} humans per GPU, MPC horizon {K}, 2 ORCA passes + integration + rewards/obs + localisation + 1 MPC solve per env-step",
            "C4": f"C4 ORCA-only step: {E} envs x {Nh}-human dense crowd per GPU (circle radius 8 m), neighbor_dist {a.neighbor_dist:g}, one ORCA pass + integration per env-step",
            "C5": f"C5 MPC solve sweep: {E} problems per GPU, horizon {K}, iteration0 solve then warm re-solves after applying the first control"}[a.config]
    c = {"workload": what, "config": a.config, "envs_per_gpu": E, "humans": Nh, "horizon": K, "scaling": a.scaling,
         "l2": "256 MB buffer written between timed steps (L2 flush)", "parallelism": "env-sharded, no step-path collective"}
    if a.config in ("C2", "C3"):
        c["policy"] = "MPC action + fixed per-env Gaussian perturbation (sigma 0.1 m/s, 0.2 rad/s), clipped to the action box"
        c["soft_reset_on_device"] = bool(a.soft_reset)
        c["done_envs"] = ("with soft_reset_on_device off (default) an env whose step terminated keeps receiving the policy's action: "
                          "same kernels, the reference would be driving it through its scripted recovery instead (--soft-reset turns that on)")
    if a.config == "C4":
        c["neighbor_dist"] = a.neighbor_dist
    return c


def _paths():
    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_paths", os.path.join(REPO, "dr-mpc_b200", "paths.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    return P.pack_paths(P.continuous_task_paths())


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/env_ref.c + rvo2_ref.c + mpc_ref.c), OpenMP over envs / problems
# ------------------------------------------------------------------------------------------------------------
def cpu_runner(a, n, threads):
    """-> (step_fn, units_per_step, description): one bounded-sample step of the config's hot path on the CPU oracle."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    tab, lens = _paths()
    if a.config in ("C2", "C3"):
        env = O.EnvRef(n, a.humans, tab, lens, K=a.horizon, soft_reset=bool(a.soft_reset))
        st = {"arr": env.reset(nthreads=threads)}
        st["act"] = st["arr"]["mpc_actions"][:, :2].copy()

        def step():
            st["arr"] = env.step(st["act"], nthreads=threads)
            st["act"] = st["arr"]["mpc_actions"][:, :2].copy()
        return step, n, f"{n} envs x {a.humans} humans, full env step, oracle/env_ref.c"
    if a.config == "C4":
        env = O.EnvRef(n, a.humans, tab, lens, neighbor_dist=a.neighbor_dist, circle_radius_humans=max(5.0, 0.16 * a.humans))
        env.reset(nthreads=threads)
        return (lambda: env.crowd_step(nthreads=threads)), n, f"{n} envs x {a.humans} humans, ORCA-only step, oracle/env_ref.c"
    K = a.horizon
    rng = np.random.default_rng(5)
    node = rng.integers(0, lens[0] - 1, n)
    st = {"pose": tab[0][:, node].T + rng.uniform(-0.3, 0.3, (n, 3)), "idx": node.astype(np.float64)}
    mb = O.MpcBatch(n, O.mpc_params(K=K), tab, lens)
    pid = np.zeros(n, np.int32)

    def step():
        u, _ = mb.act(pid, st["idx"], st["pose"], nthreads=threads)
        st["pose"], st["idx"] = _advance_np(st["pose"], st["idx"], u, lens[0])
    return step, n, f"{n} problems, horizon {K}, oracle/mpc_ref.c"


def _advance_np(pose, idx, u, L):
    import numpy as np
    v, w = u[:, 0], u[:, 1]
    w_ = np.where(w == 0, 1e-12, w)
    dx = (v / w_) * np.sin(w_ * 0.25); dy = (v / w_) * (1 - np.cos(w_ * 0.25))
    c, s = np.cos(pose[:, 2]), np.sin(pose[:, 2])
    return pose + np.stack([c * dx - s * dy, s * dx + c * dy, w * 0.25], 1), np.minimum(idx + v * 0.25 / 0.2, L - 1)


def cpu_units_per_sec(a, n, budget_s, threads):
    step, units, desc = cpu_runner(a, n, threads)
    for _ in range(2):
        step()
    t0 = time.perf_counter(); step(); one = time.perf_counter() - t0
    steps = max(3, min(2000, int(budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return units * steps / dt, steps, dt, desc


def cpu_c1_single_env(budget_s=2.0):
    """BASELINE config 1: the default scenario's shape (1 env, 6 humans, K = 4) on ONE host core, oracle port."""
    from oracle import oracle as O
    O.build()
    tab, lens = _paths()
    env = O.EnvRef(1, 6, tab, lens, K=4)
    arr = env.reset(nthreads=1)
    act = arr["mpc_actions"][:, :2].copy()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        arr = env.step(act, nthreads=1); act = arr["mpc_actions"][:, :2].copy()
        n += 1
    return n / (time.perf_counter() - t0)


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    metric, unit = METRICS[a.config]
    n = a.cpu_envs
    steps, warm = min(a.steps, 100), min(a.warmup, 5)
    step, units, desc = cpu_runner(a, n, threads)
    for _ in range(warm):
        step()
    t1 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t1
    value = units * steps / dt
    sample = f"{desc}: a {n}-unit sample of the {a.envs_per_gpu}-unit shard per step, {steps} steps after {warm} warm-up"
    cfg = workload_config(a)
    cfg["reference_arm"] = (f"bounded sample: each step is {n} units (the GPU arm's shard is {a.envs_per_gpu}), steps capped at 100 and warm-up at 5; "
                            "value is a RATE (units/s), ms_per_step is per sample step")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "ms_per_step_is_for_units": n, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = oracle/env_ref.c (+rvo2_ref.c, mpc_ref.c), the CPU restatement of the reference's RVO2+pysteam "
                    "path pinned to golden vectors of the reference's own Python; Python-RVO2 / pysteam / pylgmath are not "
                    "vendored by the reference and absent here, and /root/reference does not exist on the GPU box",
            "wall_s": time.perf_counter() - t0}
    _emit(line)


# ------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# parity gates: the config's own device state, at its own shape, through the CPU oracle
# ------------------------------------------------------------------------------------------------------------
def _copy_state_to_oracle(env, ref, n):
    for k in ("hpos", "hgoal", "hhist", "hvel", "hstatic", "hvalid", "rpose", "rprev", "rhist", "rpastv", "rvalid", "gtime",
              "path_idx", "loc_to_start", "step_count", "sr_state"):
        ref.a[k][...] = env.state[k][:n].cpu().numpy().reshape(ref.a[k].shape)
    ref.a["reset_epoch"][...] = 1
    T, u, it0 = env.mpc_get_warm()
    ref.a["warm_T"][...] = T[:n]; ref.a["warm_u"][...] = u[:n]; ref.a["iteration0"][...] = it0[:n]
    ref.a["done_bits"][...] = env.out["done_bits"][:n].cpu().numpy().view("u4")


def parity_gate_env(a, env, step_fn, rank_offset):
    """One full env step of the first n envs of the running batch on the oracle vs what the device step produces."""
    import numpy as np
    import torch
    from oracle import oracle as O
    O.build()
    tab, lens = _paths()
    n = min(env.E, 1024)
    ref = O.EnvRef(n, a.humans, tab, lens, K=a.horizon, env_id_offset=rank_offset, soft_reset=bool(a.soft_reset))
    torch.cuda.synchronize()
    _copy_state_to_oracle(env, ref, n)
    act = step_fn()                                    # the device step of the whole batch (returns the actions it used)
    torch.cuda.synchronize()
    r = ref.step(act[:n].cpu().numpy(), nthreads=0)
    st = {k: v[:n].cpu().numpy() for k, v in env.state.items()}
    o = {k: v[:n].cpu().numpy() for k, v in env.out.items()}
    stat = o["mpc_status"]
    same = (stat[:, 0] == r["mpc_status"]["iterations"]) & (stat[:, 2] == r["mpc_status"]["lm_solves"]) & (stat[:, 3] == r["mpc_status"]["cost_evals"])
    du = np.abs(o["mpc_actions"] - r["mpc_actions"]).max(1)
    return {"checked": f"{n} envs x {a.humans} humans, one full step from the running batch's own state, device vs oracle/env_ref.c",
            "orca_velocity_mismatches": int((st["hvel"].view(np.uint32) != r["hvel"].view(np.uint32)).any(-1).sum()), "orca_solves": 2 * n * a.humans,
            "max_position_err_m": float(np.abs(st["hpos"] - r["hpos"]).max()), "done_word_mismatches": int((o["done_bits"].view(np.uint32) != r["done_bits"]).sum()),
            "max_reward_err": float(np.abs(o["rewards"] - r["rewards"]).max()), "max_obs_err": float(np.abs(o["spatial_edges"] - r["spatial_edges"].reshape(o["spatial_edges"].shape)).max()),
            "lm_branch_divergences": int((~same).sum()), "mpc_solves": n, "max_control_err_same_branch": float(du[same].max()) if same.any() else None,
            "tolerance": "velocities bit-exact; positions 1e-12 m; rewards / observations 1e-9; controls 1e-6 on the same LM branch"}


def parity_gate_crowd(a, env, rank_offset):
    import numpy as np
    import torch
    import dr_mpc_b200 as d
    from oracle import oracle as O
    O.build()
    tab, lens = _paths()
    n = min(env.E, 1024)
    ref = O.EnvRef(n, a.humans, tab, lens, neighbor_dist=a.neighbor_dist, circle_radius_humans=max(5.0, 0.16 * a.humans), env_id_offset=rank_offset)
    torch.cuda.synchronize()
    _copy_state_to_oracle(env, ref, n)
    # LP-branch statistics of the very solves the step is about to do (stand-alone pass: same device solve routine)
    orca = d.BatchedORCA(env.cfg)
    orca.predict(env.state["hpos"][:n], env.state["hvel"][:n], env.state["hgoal"][:n], env.state["rpose"][:n, :2].contiguous(), want_stats=True)
    torch.cuda.synchronize()
    gs = orca.last_stats.cpu().numpy().reshape(n, a.humans, 6)
    _, os_ = O.orca_crowd_pass(ref.a["hpos"], ref.a["hvel"], ref.a["hgoal"], ref.a["rpose"][:, :2], neighbor_dist=a.neighbor_dist, want_stats=True)
    branch_div = int(((gs[..., 0] != os_["n_lines"]) | (gs[..., 1] != os_["lp2_fail"]) | (gs[..., 2] != os_["n_lp1"]) |
                      (gs[..., 4].astype(np.uint32) != os_["branch_mask"]) | (gs[..., 5] != os_["n_lp3_proj"])).sum())
    env.crowd_step()
    torch.cuda.synchronize()
    r, _ = ref.crowd_step(nthreads=0)
    st = {k: v[:n].cpu().numpy() for k, v in env.state.items()}
    return {"checked": f"{n} envs x {a.humans} humans, one ORCA-only step from the running batch's own state, device vs oracle/env_ref.c",
            "orca_velocity_mismatches": int((st["hvel"].view(np.uint32) != r["hvel"].view(np.uint32)).any(-1).sum()), "orca_solves": n * a.humans,
            "lp_branch_divergences": branch_div, "max_position_err_m": float(np.abs(st["hpos"] - r["hpos"]).max()),
            "mean_neighbours": float(os_["n_lines"].mean()), "lp3_fraction": float((os_["lp2_fail"] < os_["n_lines"]).mean()),
            "tolerance": "velocities bit-exact (fp32, FMA off); positions 1e-12 m"}


def parity_gate_mpc(a, mpc, pid, idx, pose):
    import numpy as np
    from oracle import oracle as O
    O.build()
    tab, lens = _paths()
    n = min(mpc.N, 10240)
    ob = O.MpcBatch(n, O.mpc_params(K=a.horizon), tab, lens)
    T, u, it0 = mpc.get_warm()
    ob.warm_T[:] = T[:n]; ob.warm_u[:] = u[:n]; ob.iteration0[:] = it0[:n]
    u_o, st_o = ob.act(pid[:n].cpu().numpy(), idx[:n].cpu().numpy(), pose[:n].cpu().numpy())
    return n, u_o, st_o


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return world, rank, local, dev, barrier, max_over_ranks


def _timed(step, n, flush, barrier, max_over_ranks, rank, local, sample_clocks=True):
    """Exactly n steps, per-step CUDA events on the launching stream, L2 flushed between steps -> (ms total, clocks, wall s)."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    sampler = ClockSampler(local)
    if rank == 0 and sample_clocks:
        sampler.start()
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in ev:
        s0.record(); step(); s1.record()
        flush.zero_()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    ms = max_over_ranks(sum(s0.elapsed_time(s1) for s0, s1 in ev))
    return ms, clocks, t_wall


def _peaks():
    from dr_mpc_b200 import _lib
    p32, p64 = _lib.c_d(), _lib.c_d()
    _lib.check(_lib.load().dre_measure_peak(0, p32)); _lib.check(_lib.load().dre_measure_peak(1, p64))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return p32.value, p64.value, hbm, src


def _counted():
    """Instrumented op counts of the flat oracle (oracle/opcount/count.py -> profiles/roofline.json)."""
    try:
        return json.load(open(os.path.join(REPO, "profiles", "roofline.json")))
    except Exception:
        return None


def _orca_flops(counted, Nh, stats):
    """flops per ORCA solve: the instrumented count of the shape when there is one (scaled by this state's LP1 work through
    the survey's formula), else the survey's formula 53 n + sum over LP1 calls (10 i + 20)."""
    formula = 53.0 * Nh + 10.0 * float(stats[..., 3].double().mean()) + 20.0 * float(stats[..., 2].double().mean())
    c = (counted or {}).get("orca", {}).get(f"n{Nh}")
    if c:
        return c["flops_per_solve"] * formula / c["survey_formula_53n_plus_lp1"], "instrumented", formula
    return formula, "survey formula", formula


def _mpc_flops(counted, K):
    c = (counted or {}).get("mpc", {}).get(f"K{K}")
    if c:
        return c["flops_per_damped_solve"], "instrumented"
    return 2000.0 * K, "survey estimate"


def run_env(a):
    """C2 / C3: the full env step."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import dr_mpc_b200 as d
    world, rank, local, dev, barrier, max_over_ranks = _dist_setup()
    E, Nh, K = a.envs_per_gpu, a.humans, a.horizon
    metric, unit = METRICS[a.config]
    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, horizon=K, env_id_offset=rank * E, soft_reset_on_device=a.soft_reset), num_envs=E, device=dev)
    env.reset()
    act = env.out["mpc_actions"][:, :2].clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches_per_step = 4                      # pt_step, ha_step, mpc_solve, mpc_stats

    # policy: the MPC action plus a fixed per-env perturbation (N(0, 0.1) on v, N(0, 0.2) on w, clipped to the action box).
    # Without it every robot would follow the identical trajectory and all E MPC problems would be the same one.
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise = torch.randn((64, E, 2), generator=gen, device=dev, dtype=torch.float64) * torch.tensor([0.1, 0.2], device=dev, dtype=torch.float64)
    lo = torch.tensor([0.0, -1.0], device=dev, dtype=torch.float64)
    hi = torch.tensor([1.0, 1.0], device=dev, dtype=torch.float64)
    tick = [0]

    def step():
        torch.add(env.out["mpc_actions"][:, :2], noise[tick[0] % 64], out=act)
        torch.clamp(act, lo, hi, out=act)
        tick[0] += 1
        env.step(act)
        return act

    for _ in range(max(a.warmup, 3)):
        step(); flush.zero_()
    parity = None
    if rank == 0 and not a.no_parity_gate:
        parity = parity_gate_env(a, env, step, rank * E)
    env.stats(reset=True)
    total_ms, clocks, t_wall = _timed(step, a.steps, flush, barrier, max_over_ranks, rank, local)
    value = world * E * a.steps / (total_ms * 1e-3)
    # the only collective of the job: the episode statistics, reduced by the engine itself (dre_stats_reduce over NCCL / NVLink)
    stats_via = "single process"
    if world > 1:
        try:
            d.parallel.init_stats_comm(env)
            cnt, sums = env.stats_reduce()
            stats_via = "dre_stats_reduce (ncclAllReduce issued by libdre)"
        except Exception as ex:   # noqa: BLE001 -- keep the bench line; say what happened
            cnt, sums = env.stats()
            dist.all_reduce(cnt); dist.all_reduce(sums)
            stats_via = f"torch.distributed all_reduce (dre_stats_reduce unavailable: {ex})"
    else:
        cnt, sums = env.stats()
    cnt = cnt.cpu().numpy(); sums = sums.cpu().numpy()
    if int(cnt[0]) != world * E * a.steps:
        raise SystemExit(f"episode statistics do not add up: {int(cnt[0])} env-steps counted, {world * E * a.steps} timed ({stats_via})")

    # ---------------- per-kernel durations (same loop, events around every kernel) ----------------
    env.set_profiling(True)
    env.kernel_ms(reset=True)
    for _ in range(min(a.steps, 100)):
        step(); flush.zero_()
    kms, nprof = env.kernel_ms(reset=True)
    env.set_profiling(False)
    kavg = {k: v / max(nprof, 1) for k, v in kms.items()}
    counted = _counted()
    orca = d.BatchedORCA(env.cfg)
    orca.predict(env.state["hpos"], env.state["hvel"], env.state["hgoal"], env.state["rpose"][:, :2].contiguous(), want_stats=True)
    torch.cuda.synchronize()
    flops_solve, orca_src, formula = _orca_flops(counted, Nh, orca.last_stats)
    ha_flops = 2 * E * Nh * flops_solve
    ha_bytes = 2 * E * (40 * Nh + 32)
    mst = env.out["mpc_status"].double()
    mpc_per_solve, mpc_src = _mpc_flops(counted, K)
    mpc_flops = E * mpc_per_solve * float(mst[:, 2].mean())
    mpc_bytes = E * (15 * K + 3) * 8
    peak32, peak64, hbm_peak, hbm_src = _peaks()
    dom = max(("ha_step", "mpc_solve", "pt_step"), key=lambda k: kavg[k])
    dom_bytes = ha_bytes if dom == "ha_step" else (mpc_bytes if dom == "mpc_solve" else E * (4 * 8 * 3 + 800))
    dom_flops, dom_peak, dom_kind = (ha_flops, peak32, "fp32 CUDA-core") if dom == "ha_step" else (mpc_flops, peak64, "fp64 CUDA-core")
    traffic = issue = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get(dom)
        issue = json.load(open(os.path.join(REPO, "profiles", "issue.json")))
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": dom_bytes / (kavg[dom] * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "frac": dom_bytes / (kavg[dom] * 1e-3) * 1e-9 / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                "note": "the path is CUDA-core latency/issue bound, not HBM- or tensor-bound (SURVEY 8(d)); see roofline_compute"}
    roofline_compute = {"kernel": dom, "bound": dom_kind, "achieved": dom_flops / (kavg[dom] * 1e-3) * 1e-12, "peak": dom_peak,
                        "unit": "TFLOP/s", "frac": dom_flops / (kavg[dom] * 1e-3) * 1e-12 / dom_peak,
                        "peak_source": "FMA-chain microbenchmark run in this process (dre_measure_peak)",
                        "algorithmic_flops_per_launch": dom_flops,
                        "all": {"ha_step": {"tflops": ha_flops / (kavg["ha_step"] * 1e-3) * 1e-12, "frac_fp32": ha_flops / (kavg["ha_step"] * 1e-3) * 1e-12 / peak32,
                                            "flops_per_orca_solve": flops_solve, "flops_source": orca_src, "survey_formula": formula},
                                "mpc_solve": {"tflops": mpc_flops / (kavg["mpc_solve"] * 1e-3) * 1e-12, "frac_fp64": mpc_flops / (kavg["mpc_solve"] * 1e-3) * 1e-12 / peak64,
                                              "flops_per_damped_solve": mpc_per_solve, "flops_source": mpc_src,
                                              "mean_lm_solves": float(mst[:, 2].mean()), "mean_lm_iters": float(mst[:, 0].mean())}},
                        "flops_source": orca_src if dom == "ha_step" else mpc_src,
                        "peaks_tflops": {"fp32": peak32, "fp64": peak64},
                        "ncu_issue_slots": issue,
                        "note": "algorithmic flops are COUNTED on the flat oracle compiled with instrumented scalars (oracle/opcount, profiles/roofline.json); "
                                "they are a small part of the issued instructions: the kernels are bound by instruction issue (ha_step) and by dependent fp64 "
                                "chains / instruction issue and fetch (mpc_solve, profiles/README.md)"}

    # ---------------- de-duplicated variant, reported separately: the look-ahead ORCA pass of step t IS the
    # movement pass of step t+1 unless a goal was resampled in between; bit-identical results, one pass per step ----------------
    def variant(**kw):
        env2 = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, horizon=K, env_id_offset=rank * E, soft_reset_on_device=a.soft_reset, **kw), num_envs=E, device=dev)
        env2.reset()
        act2 = env2.out["mpc_actions"][:, :2].clone()
        tk = [0]

        def step2():
            torch.add(env2.out["mpc_actions"][:, :2], noise[tk[0] % 64], out=act2)
            torch.clamp(act2, lo, hi, out=act2)
            tk[0] += 1
            env2.step(act2)

        for _ in range(max(a.warmup, 3)):
            step2(); flush.zero_()
        ms2, _, _ = _timed(step2, a.steps, flush, barrier, max_over_ranks, rank, local, sample_clocks=False)
        del env2
        return {"value": world * E * a.steps / (ms2 * 1e-3), "unit": unit, "ms_per_step": ms2 / a.steps}

    dedup = pruned = None
    if not a.no_dedup:
        dedup = variant(orca_dedup=True)
        dedup["what"] = ("EngineConfig(orca_dedup=True): same steps, same policy, same results bit for bit (tests/test_gpu_env.py::"
                         "test_orca_dedup_is_bit_identical); NOT the headline, which keeps the reference's two ORCA passes per step")
        pruned = variant(orca_lookahead_prune=True)
        pruned["what"] = ("EngineConfig(orca_lookahead_prune=True): the look-ahead pass solves only the humans whose future velocity the "
                          "disturbance reward reads (within 1.5 m of the robot); same results bit for bit (tests/test_gpu_env.py::"
                          "test_lookahead_prune_is_bit_identical); NOT the headline, which keeps the reference's two full ORCA passes per step")

    # ---------------- the data-collection loop's other half, reported separately: every step followed by the replay insert
    # (online_continuous_task.py:185-189) into the device-resident ring (dre_replay_insert: select + gather, HBM-bound) ----------------
    replay = None
    if not a.no_dedup:
        try:
            rb = d.DeviceReplayBuffer(env, 262144)

            def step_insert():
                rb.insert(step())

            for _ in range(3):
                step_insert(); flush.zero_()
            ms3, _, _ = _timed(step_insert, a.steps, flush, barrier, max_over_ranks, rank, local, sample_clocks=False)
            ms_ins, _, _ = _timed(lambda: rb.insert(act), min(a.steps, 20), flush, barrier, max_over_ranks, rank, local, sample_clocks=False)
            ms_ins /= min(a.steps, 20)
            W = sum(rb._widths)
            ins_bytes = E * (2 * W * 8 + 29) + E * (2 * W + 4) * 8
            replay = {"value": world * E * a.steps / (ms3 * 1e-3), "unit": unit, "ms_per_step": ms3 / a.steps, "insert_ms": ms_ins,
                      "insert_algorithmic_bytes": ins_bytes, "insert_GBps": ins_bytes / (ms_ins * 1e-3) * 1e-9,
                      "insert_frac_of_hbm_peak": ins_bytes / (ms_ins * 1e-3) * 1e-9 / hbm_peak, "ring_rows": rb.buffer_size,
                      "ring_bytes": rb.storage_bytes, "valid_entries": rb.valid_entries,
                      "what": "every step followed by dre_replay_insert of all E transitions (S from the previous slab, S' from the latest, "
                              "float64 ring with the reference ReplayBuffer's layout); NOT the headline, which times the env step alone. Timed later in "
                              "the episode than the headline window (sparser crowd, faster steps): compare insert_ms, not ms_per_step"}
            del rb
        except Exception as ex:   # noqa: BLE001 -- an extra; never lose the bench line over it
            replay = {"error": repr(ex)}

    # ---------------- end to end through the host-buffer C-ABI call (closed loop through host memory) ----------------
    noise_h = noise[:8].cpu().numpy()
    a_host = torch.empty((E, 2), dtype=torch.float64).pin_memory().numpy()
    a_host[:] = env.out["mpc_actions"][:, :2].cpu().numpy()
    noise_c = np.ascontiguousarray(noise_h.transpose(0, 2, 1))   # [8][2][E]: the policy works column by column
    a0_h, a1_h = a_host[:, 0], a_host[:, 1]

    def host_policy(views, i):
        ma = views["mpc_actions"]
        np.add(ma[:, 0], noise_c[i % 8, 0], out=a0_h); np.clip(a0_h, 0.0, 1.0, out=a0_h)
        np.add(ma[:, 1], noise_c[i % 8, 1], out=a1_h); np.clip(a1_h, -1.0, 1.0, out=a1_h)

    def e2e_loop(step_host, n):
        for i in range(3):
            host_policy(step_host(a_host), i)
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            host_policy(step_host(a_host), i)   # the next action is read from the host slab
        torch.cuda.synchronize()
        s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return s

    n_e2e = a.steps
    e2e_s = e2e_loop(env.step_host, n_e2e)
    e2e = {"value": world * E * n_e2e / e2e_s, "unit": unit, "h2d_bytes_per_step": E * 2 * 8, "d2h_bytes_per_step": env.slab_bytes,
           "steps": n_e2e, "ms_per_step": e2e_s / n_e2e * 1e3,
           "what": "dre_env_step_host: pinned host actions -> H2D -> robot/path half (its PT state travels during the MPC solve) -> MPC "
                   "-> one crowd-half launch whose env chunks raise flags in mapped host memory as their observations are written; "
                   "the host thread queues each chunk's D2H on a copy stream -> the whole output slab (observations, rewards, done "
                   "bits, MPC status) in pinned host memory; the next action is read from the host slab"}
    # separately reported: (a) the D2H ceiling of this box for the same bytes (plain pinned copies, all ranks at once), (b) the
    # slab's observation part as float32 (exact single rounding of the float64 results), (c) only what a caller whose networks
    # sit on the GPU needs on the host (actions / rewards / done words; the reference moves S to cuda:0 anyway, misc.py:7-21)
    e2e_extra = {}
    if hasattr(env, "d2h_ceiling_ms"):
        ceil_ms = max_over_ranks(env.d2h_ceiling_ms(reps=20))
        e2e_extra["d2h_ceiling"] = {"ms_per_slab": ceil_ms, "gb_s_per_rank": env.slab_bytes / (ceil_ms * 1e-3) * 1e-9,
                                    "e2e_fraction_of_ceiling": ceil_ms / (e2e_s / n_e2e * 1e3),
                                    "what": "one plain pinned cudaMemcpyAsync D2H of slab_bytes per rank, all ranks concurrently, max over ranks: "
                                            "the floor of a step that returns the whole float64 slab on this box"}
    for name, kw, what in (("f32_observations", {"out_dtype": "float32"}, "observation arrays rounded once to float32 on the device before the copy (rewards / done words / MPC actions stay float64)"),
                           ("control_only", {"outputs": "control"}, "only MPC actions, executed actions, rewards, info, done words and MPC status travel; the observations stay on the GPU")):
        if hasattr(env, "configure_host_outputs"):
            env.configure_host_outputs(**kw)
            s = e2e_loop(env.step_host, n_e2e)
            e2e_extra[name] = {"value": world * E * n_e2e / s, "unit": unit, "ms_per_step": s / n_e2e * 1e3, "d2h_bytes_per_step": env.host_bytes_per_step, "what": what}
            env.configure_host_outputs()
    if e2e_extra:
        e2e["variants"] = e2e_extra

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, nst, dt, desc = cpu_units_per_sec(a, min(a.cpu_envs, E), 12.0, threads)
        cpu_baseline = {"value": v, "unit": unit, "cores": threads, "kind": "port",
                        "sample": f"{desc}: {min(a.cpu_envs, E)} of the {E} envs, {nst} steps ({dt:.1f} s), OpenMP over envs",
                        "c1_single_env_1core": {"value": cpu_c1_single_env(), "unit": unit,
                                                "what": "BASELINE configs[0] shape: 1 env x 6 humans, K = 4, one host core, same oracle port "
                                                        "(the reference's own Python on the API shims does ~43 env-steps/s on this shape)"}}
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": workload_config(a), "clocks": clocks,
                "e2e": e2e, "gpu_launches": launches_per_step * a.steps, "parity": parity, "roofline": roofline, "roofline_compute": roofline_compute,
                "kernels_ms": kavg, "dedup_variant": dedup, "lookahead_prune_variant": pruned, "replay_variant": replay, "cpu_baseline": cpu_baseline, "wall_ms_per_step_incl_flush": t_wall / a.steps * 1e3,
                "episode_stats": {n: int(cnt[i]) for i, n in enumerate(d.env.STAT_NAMES)}, "episode_stats_via": stats_via,
                "reward_sums": [float(x) for x in sums[:3]]}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_crowd(a):
    """C4: ORCA-only step of a dense crowd."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import dr_mpc_b200 as d
    world, rank, local, dev, barrier, max_over_ranks = _dist_setup()
    E, Nh = a.envs_per_gpu, a.humans
    metric, unit = METRICS[a.config]
    cfg = d.EngineConfig(num_humans=Nh, horizon=a.horizon, env_id_offset=rank * E, neighbor_dist=a.neighbor_dist, circle_radius_humans=max(5.0, 0.16 * Nh))
    env = d.BatchedHAAndPTEnv(cfg, num_envs=E, device=dev)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(a.warmup, 3)):
        env.crowd_step(); flush.zero_()
    parity = parity_gate_crowd(a, env, rank * E) if (rank == 0 and not a.no_parity_gate) else None
    total_ms, clocks, t_wall = _timed(env.crowd_step, a.steps, flush, barrier, max_over_ranks, rank, local)
    value = world * E * a.steps / (total_ms * 1e-3)
    counted = _counted()
    orca = d.BatchedORCA(env.cfg)
    orca.predict(env.state["hpos"], env.state["hvel"], env.state["hgoal"], env.state["rpose"][:, :2].contiguous(), want_stats=True)
    torch.cuda.synchronize()
    flops_solve, src, formula = _orca_flops(counted, Nh, orca.last_stats)
    n_nb = float(orca.last_stats[..., 0].double().mean())
    if a.neighbor_dist < 10.0:                      # only the in-range neighbours are algorithmic work
        flops_solve, src = 53.0 * n_nb + 10.0 * float(orca.last_stats[..., 3].double().mean()) + 20.0 * float(orca.last_stats[..., 2].double().mean()), "survey formula on the in-range neighbours"
    flops = E * Nh * flops_solve
    ms = total_ms / a.steps
    peak32, peak64, hbm_peak, hbm_src = _peaks()
    nbytes = E * (40 * Nh + 32)
    roofline = {"kernel": "ha_step (warm-start mode: one ORCA pass + integration)", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) * 1e-9, "peak": hbm_peak,
                "unit": "GB/s", "frac": nbytes / (ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None, "peak_source": hbm_src,
                "note": "CUDA-core issue bound, not HBM bound; see roofline_compute"}
    roofline_compute = {"kernel": "ha_step (warm-start mode)", "bound": "fp32 CUDA-core", "achieved": flops / (ms * 1e-3) * 1e-12, "peak": peak32, "unit": "TFLOP/s",
                        "frac": flops / (ms * 1e-3) * 1e-12 / peak32, "flops_per_orca_solve": flops_solve, "flops_source": src,
                        "mean_neighbours_in_range": n_nb, "peak_source": "FMA-chain microbenchmark run in this process (dre_measure_peak)"}
    # e2e: the crowd state a host-side caller needs back each step (positions + velocities), no inputs
    hp = torch.empty((E, Nh, 2), dtype=torch.float64).pin_memory(); hv = torch.empty((E, Nh, 2), dtype=torch.float32).pin_memory()

    def host_step():
        env.crowd_step()
        hp.copy_(env.state["hpos"], non_blocking=True); hv.copy_(env.state["hvel"], non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(3):
        host_step()
    barrier(); t0 = time.perf_counter()
    for _ in range(a.steps):
        host_step()
    e2e_s = max_over_ranks(time.perf_counter() - t0); barrier()
    e2e = {"value": world * E * a.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": E * Nh * 24, "ms_per_step": e2e_s / a.steps * 1e3,
           "what": "dre_env_crowd_step + D2H of the humans' positions (f64) and velocities (f32) into pinned host memory every step"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, nst, dt, desc = cpu_units_per_sec(a, min(a.cpu_envs, E), 12.0, threads)
        cpu_baseline = {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": f"{desc}, {nst} steps ({dt:.1f} s)"}
    if rank == 0:
        _emit({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
               "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32 ORCA + f64 positions", "data": "synthetic",
               "config": workload_config(a), "clocks": clocks, "e2e": e2e, "gpu_launches": a.steps, "parity": parity, "roofline": roofline,
               "roofline_compute": roofline_compute, "cpu_baseline": cpu_baseline, "orca_solves_per_sec": value * Nh})
    if world > 1:
        dist.destroy_process_group()


def run_mpc(a):
    """C5: MPC solve sweep (one horizon per invocation)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import dr_mpc_b200 as d
    world, rank, local, dev, barrier, max_over_ranks = _dist_setup()
    N, K = a.envs_per_gpu, a.horizon
    metric, unit = METRICS[a.config]
    tab, lens = _paths()
    paths = d.paths.continuous_task_paths()
    m = d.BatchedMPC(None, K, 0.25, True, 1.0, 1.0, num_problems=N)
    m.set_paths(paths)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    L0 = int(lens[0])
    node = torch.randint(0, L0 - 1, (N,), generator=gen, device=dev)
    p0 = torch.tensor(paths[0], device=dev)
    pose = p0[:, node].T.contiguous() + (torch.rand((N, 3), generator=gen, device=dev, dtype=torch.float64) - 0.5) * 0.6
    idx = node.double()
    pid = torch.zeros(N, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    state = {"pose": pose, "idx": idx, "u": None}

    def advance(u):
        v, w = u[:, 0], u[:, 1]
        w_ = torch.where(w == 0, torch.full_like(w, 1e-12), w)
        dx = (v / w_) * torch.sin(w_ * 0.25); dy = (v / w_) * (1 - torch.cos(w_ * 0.25))
        c, s = torch.cos(state["pose"][:, 2]), torch.sin(state["pose"][:, 2])
        state["pose"] = state["pose"] + torch.stack([c * dx - s * dy, s * dx + c * dy, w * 0.25], 1)
        state["idx"] = torch.clamp(state["idx"] + v * 0.25 / 0.2, max=L0 - 1)

    # iteration0 solve (timed on its own), then warm re-solves; the state advance between solves is outside the timed events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    u = m.act({'path_id': pid, 'interp_index': state["idx"], 'pose': state["pose"]}, want_status=True)
    e1.record(); barrier()
    it0_ms = max_over_ranks(e0.elapsed_time(e1))
    it0_solves = float(m.last_status[:, 2].double().mean())
    advance(u)
    for _ in range(max(a.warmup, 3) - 1):
        u = m.act({'path_id': pid, 'interp_index': state["idx"], 'pose': state["pose"]}); advance(u); flush.zero_()
    parity = None
    if rank == 0 and not a.no_parity_gate:
        n, u_o, st_o = parity_gate_mpc(a, m, pid, state["idx"], state["pose"])
        u = m.act({'path_id': pid, 'interp_index': state["idx"], 'pose': state["pose"]}, want_status=True)
        torch.cuda.synchronize()
        st = m.last_status[:n].cpu().numpy()
        same = (st[:, 0] == st_o["iterations"]) & (st[:, 2] == st_o["lm_solves"]) & (st[:, 3] == st_o["cost_evals"]) & (st[:, 1] == st_o["retries"])
        du = np.abs(u[:n].cpu().numpy() - u_o).max(1)
        parity = {"checked": f"{n} problems, horizon {K}, one warm re-solve from the running batch's own warm start, device vs oracle/mpc_ref.c",
                  "lm_branch_divergences": int((~same).sum()), "mpc_solves": n, "max_control_err_same_branch": float(du[same].max()),
                  "fraction_within_1e-6": float((du[same] <= 1e-6).mean()), "tolerance": "controls 1e-6 (5e-6 for K > 10) on the same LM branch"}
        advance(u)
    elif not a.no_parity_gate:
        u = m.act({'path_id': pid, 'interp_index': state["idx"], 'pose': state["pose"]}); advance(u)
    solves = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    for s0, s1 in ev:
        s0.record()
        u = m.act({'path_id': pid, 'interp_index': state["idx"], 'pose': state["pose"]}, want_status=True)
        s1.record()
        solves.append(m.last_status[:, 2].double().mean())
        advance(u); flush.zero_()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(sum(s0.elapsed_time(s1) for s0, s1 in ev))
    ms = total_ms / a.steps
    value = world * N * a.steps / (total_ms * 1e-3)
    mean_solves = float(torch.stack(solves).mean())
    per_solve, src = _mpc_flops(_counted(), K)
    peak32, peak64, hbm_peak, hbm_src = _peaks()
    flops = N * per_solve * mean_solves
    nbytes = N * (15 * K + 3) * 8
    roofline = {"kernel": "mpc_coop_kernel", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "frac": nbytes / (ms * 1e-3) * 1e-9 / hbm_peak, "traffic": None, "peak_source": hbm_src, "note": "fp64 CUDA-core / latency bound; see roofline_compute"}
    roofline_compute = {"kernel": "mpc_coop_kernel", "bound": "fp64 CUDA-core", "achieved": flops / (ms * 1e-3) * 1e-12, "peak": peak64, "unit": "TFLOP/s",
                        "frac": flops / (ms * 1e-3) * 1e-12 / peak64, "flops_per_damped_solve": per_solve, "flops_source": src, "mean_lm_solves": mean_solves,
                        "peak_source": "FMA-chain microbenchmark run in this process (dre_measure_peak)"}
    # e2e: dre_mpc_act_host (interp index + pose in, controls out through host memory)
    n_e2e = max(3, min(a.steps, 20))
    idx_h, pose_h, pid_h = state["idx"].cpu().numpy(), state["pose"].cpu().numpy(), pid.cpu().numpy()
    m.act({'path_id': pid_h, 'interp_index': idx_h, 'pose': pose_h})
    barrier(); t0 = time.perf_counter()
    for _ in range(n_e2e):
        u_h = m.act({'path_id': pid_h, 'interp_index': idx_h, 'pose': pose_h})
        pose_h, idx_h = _advance_np(pose_h, idx_h, u_h, L0)
    e2e_s = max_over_ranks(time.perf_counter() - t0); barrier()
    e2e = {"value": world * N * n_e2e / e2e_s, "unit": unit, "h2d_bytes_per_step": N * (4 + 8 + 24), "d2h_bytes_per_step": N * 2 * K * 8, "steps": n_e2e,
           "ms_per_step": e2e_s / n_e2e * 1e3, "what": "dre_mpc_act_host: path id, interp index, pose from host memory, controls back to host memory; the host applies the first control (numpy) between solves"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, nst, dt, desc = cpu_units_per_sec(a, min(a.cpu_envs, N), 12.0, threads)
        cpu_baseline = {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": f"{desc}, {nst} steps ({dt:.1f} s)"}
    if rank == 0:
        _emit({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
               "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a),
               "clocks": clocks, "e2e": e2e, "gpu_launches": a.steps, "parity": parity, "roofline": roofline, "roofline_compute": roofline_compute,
               "iteration0_solve": {"ms": it0_ms, "solves_per_sec": world * N / (it0_ms * 1e-3), "mean_lm_solves": it0_solves}, "cpu_baseline": cpu_baseline})
    if world > 1:
        dist.destroy_process_group()


def _emit(line):
    """The JSON line goes to the process's real stdout; everything else written to fd 1 (NCCL's version banner, library
    chatter) was pointed at stderr by _quiet_stdout so that stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    sys.stdout.flush()
    while data:
        data = data[os.write(fd, data):]


_REAL_STDOUT = None


def _quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.config in ("C2", "C3"):
        run_env(args)
    elif args.config == "C4":
        run_crowd(args)
    else:
        run_mpc(args)
