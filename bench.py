#!/usr/bin/env python
"""bench.py -- env-steps/sec of the DR-MPC hot path (ORCA crowd + MPC solve + env step) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU path on the box's host cores

One "step" = one full env step of every env of the shard: 2 ORCA passes over Nh humans (movement + the disturbance
look-ahead, like the reference), robot/human integration, HA rewards + observation, path localisation + PT state +
PT rewards, one MPC solve. Workload = BASELINE.json configs[2]: 16384 envs x 20 humans, K = 4, PER GPU
(weak scaling: envs are independent, every rank owns its own shard, no collective on the step path; episode
statistics are all-reduced once after the timed region).

Prints ONE JSON line (rank 0). `value` = device-timed whole-job throughput with state resident in HBM; `e2e` = the
same through the host-buffer C-ABI call (H2D of the actions + D2H of the whole step result every step, closed loop
through host memory); `roofline` / `roofline_compute` / `kernels` explain it; `cpu_baseline` = the CPU oracle port
timed on this box's host cores on a bounded sample.
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

METRIC = "env-steps/sec (ORCA crowd + MPC solve)"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=16384)
    ap.add_argument("--humans", type=int, default=20)
    ap.add_argument("--horizon", type=int, default=4)
    ap.add_argument("--cpu-envs", type=int, default=4096, help="envs of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dedup", action="store_true", help="skip the separately reported orca_dedup variant")
    return ap.parse_args()


def workload_config(a):
    return {"workload": f"C3 full env step: {a.envs_per_gpu} envs x {a.humans} humans per GPU, MPC horizon {a.horizon}, "
                        "2 ORCA passes + integration + rewards/obs + localisation + 1 MPC solve per env-step",
            "envs_per_gpu": a.envs_per_gpu, "humans": a.humans, "horizon": a.horizon, "policy": "MPC action + fixed per-env Gaussian perturbation (sigma 0.1 m/s, 0.2 rad/s), clipped to the action box",
            "l2": "256 MB buffer written between timed steps (L2 flush)", "parallelism": "env-sharded, no step-path collective"}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference env step (oracle/env_ref.c), all host threads
# ------------------------------------------------------------------------------------------------------------
def cpu_env_steps_per_sec(a, envs, budget_s, threads):
    import numpy as np
    from oracle import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_paths", os.path.join(REPO, "dr-mpc_b200", "paths.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    O.build()
    tab, lens = P.pack_paths(P.continuous_task_paths())
    env = O.EnvRef(envs, a.humans, tab, lens, K=a.horizon)
    arr = env.reset(nthreads=threads)
    act = arr["mpc_actions"][:, :2].copy()
    for _ in range(2):
        arr = env.step(act, nthreads=threads); act = arr["mpc_actions"][:, :2].copy()
    t0 = time.perf_counter()
    arr = env.step(act, nthreads=threads); act = arr["mpc_actions"][:, :2].copy()
    one = time.perf_counter() - t0
    steps = max(3, min(2000, int(budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        arr = env.step(act, nthreads=threads); act = arr["mpc_actions"][:, :2].copy()
    dt = time.perf_counter() - t0
    return envs * steps / dt, steps, dt


def cpu_c1_single_env(a, budget_s=2.0):
    """BASELINE config 1: the default scenario's shape (1 env, 6 humans, K = 4) on ONE host core, oracle port."""
    import importlib.util
    from oracle import oracle as O
    spec = importlib.util.spec_from_file_location("dre_paths", os.path.join(REPO, "dr-mpc_b200", "paths.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    O.build()
    tab, lens = P.pack_paths(P.continuous_task_paths())
    env = O.EnvRef(1, 6, tab, lens, K=4)
    arr = env.reset(nthreads=1)
    act = arr["mpc_actions"][:, :2].copy()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        arr = env.step(act, nthreads=1); act = arr["mpc_actions"][:, :2].copy()
        n += 1
    return n / (time.perf_counter() - t0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    # K "steps" of a bounded sample each; the whole run stays within a few minutes
    import numpy as np
    from oracle import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_paths", os.path.join(REPO, "dr-mpc_b200", "paths.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    O.build()
    tab, lens = P.pack_paths(P.continuous_task_paths())
    envs = a.cpu_envs
    steps = min(a.steps, 100)
    warm = min(a.warmup, 5)
    env = O.EnvRef(envs, a.humans, tab, lens, K=a.horizon)
    arr = env.reset(nthreads=threads)
    act = arr["mpc_actions"][:, :2].copy()
    for _ in range(warm):
        arr = env.step(act, nthreads=threads); act = arr["mpc_actions"][:, :2].copy()
    t1 = time.perf_counter()
    for _ in range(steps):
        arr = env.step(act, nthreads=threads); act = arr["mpc_actions"][:, :2].copy()
    dt = time.perf_counter() - t1
    value = envs * steps / dt
    sample = f"{envs} of the {a.envs_per_gpu} envs x {a.humans} humans per step, {steps} steps after {warm} warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 ORCA + f64 MPC/env",
            "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dr_mpc_b200 as d
    from dr_mpc_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, Nh, K = a.envs_per_gpu, a.humans, a.horizon

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

    env = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, horizon=K, env_id_offset=rank * E), num_envs=E, device=dev)
    S = env.reset()
    act = S['PT']['MPC_actions'][:, :2].clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches_per_step = 4                      # pt_step, mpc_solve, mpc_stats, ha_step

    # policy: the MPC action plus a fixed per-env perturbation (N(0, 0.1) on v, N(0, 0.2) on w, clipped to the action
    # box). Without it every robot would follow the identical trajectory and all E MPC problems would be the same one.
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

    for _ in range(max(a.warmup, 3)):
        step(); flush.zero_()
    env.stats(reset=True)
    # ---------------- timed region: exactly K steps, per-step CUDA events, L2 flushed between steps ----------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in ev:
        s0.record(); step(); s1.record()
        flush.zero_()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s0.elapsed_time(s1) for s0, s1 in ev)
    total_ms = max_over_ranks(dev_ms)
    value = world * E * a.steps / (total_ms * 1e-3)
    cnt, sums = env.stats()
    if world > 1:                               # the only collective of the job: episode statistics over NVLink
        dist.all_reduce(cnt); dist.all_reduce(sums)
    cnt = cnt.cpu().numpy(); sums = sums.cpu().numpy()

    # ---------------- per-kernel durations (same loop, events around every kernel) ----------------
    env.set_profiling(True)
    env.kernel_ms(reset=True)
    for _ in range(min(a.steps, 100)):
        step(); flush.zero_()
    kms, nprof = env.kernel_ms(reset=True)
    env.set_profiling(False)
    kavg = {k: v / max(nprof, 1) for k, v in kms.items()}
    # algorithmic work of the current state: ORCA flops = 53 n + sum over LP1 calls (10 i + 20); MPC = 2000 K per damped solve
    orca = d.BatchedORCA(env.cfg)
    orca.predict(env.state["hpos"], env.state["hvel"], env.state["hgoal"], env.state["rpose"][:, :2].contiguous(), want_stats=True)
    torch.cuda.synchronize()
    st = orca.last_stats.double()
    n_nb = Nh                                    # Nh-1 other humans + the robot
    flops_solve = 53.0 * n_nb + 10.0 * float(st[..., 3].mean()) + 20.0 * float(st[..., 2].mean())
    ha_flops = 2 * E * Nh * flops_solve
    ha_bytes = 2 * E * (40 * Nh + 32)
    mst = env.out["mpc_status"].double()
    mpc_flops = E * 2000.0 * K * float(mst[:, 2].mean())
    mpc_bytes = E * (15 * K + 3) * 8
    peak32, peak64 = _lib.c_d(), _lib.c_d()
    _lib.check(_lib.load().dre_measure_peak(0, peak32)); _lib.check(_lib.load().dre_measure_peak(1, peak64))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(("ha_step", "mpc_solve", "pt_step"), key=lambda k: kavg[k])
    dom_bytes = ha_bytes if dom == "ha_step" else (mpc_bytes if dom == "mpc_solve" else E * (4 * 8 * 3 + 800))
    dom_flops, dom_peak, dom_kind = (ha_flops, peak32.value, "fp32 CUDA-core") if dom == "ha_step" else (mpc_flops, peak64.value, "fp64 CUDA-core")
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    issue = None
    try:
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
                        "all": {"ha_step": {"tflops": ha_flops / (kavg["ha_step"] * 1e-3) * 1e-12, "frac_fp32": ha_flops / (kavg["ha_step"] * 1e-3) * 1e-12 / peak32.value,
                                            "flops_per_orca_solve": flops_solve},
                                "mpc_solve": {"tflops": mpc_flops / (kavg["mpc_solve"] * 1e-3) * 1e-12, "frac_fp64": mpc_flops / (kavg["mpc_solve"] * 1e-3) * 1e-12 / peak64.value,
                                              "mean_lm_solves": float(mst[:, 2].mean()), "mean_lm_iters": float(mst[:, 0].mean())}},
                        "peaks_tflops": {"fp32": peak32.value, "fp64": peak64.value},
                        "ncu_issue_slots": issue,
                        "note": "algorithmic flops (SURVEY 8(d)) are a small part of the issued instructions: the kernels are bound by "
                                "instruction issue (ha_step) and by dependent fp64 chains / instruction issue and fetch (mpc_solve, profiles/README.md); "
                                "ncu_issue_slots holds the issue-slot utilisation of the committed ncu capture"}

    # ---------------- de-duplicated variant, reported separately: the look-ahead ORCA pass of step t IS the
    # movement pass of step t+1 unless a goal was resampled in between; bit-identical results, one pass per step ----------------
    dedup = None
    if not a.no_dedup:
        env2 = d.BatchedHAAndPTEnv(d.EngineConfig(num_humans=Nh, horizon=K, env_id_offset=rank * E, orca_dedup=True), num_envs=E, device=dev)
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
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        barrier()
        for s0, s1 in ev2:
            s0.record(); step2(); s1.record()
            flush.zero_()
        barrier()
        ms2 = max_over_ranks(sum(s0.elapsed_time(s1) for s0, s1 in ev2))
        dedup = {"value": world * E * a.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / a.steps,
                 "what": "EngineConfig(orca_dedup=True): same steps, same policy, same results bit for bit (tests/test_gpu_env.py::"
                         "test_orca_dedup_is_bit_identical); NOT the headline, which keeps the reference's two ORCA passes per step"}
        del env2

    # ---------------- end to end through the host-buffer C-ABI call (closed loop through host memory) ----------------
    noise_h = noise[:8].cpu().numpy()
    a_host = torch.empty((E, 2), dtype=torch.float64).pin_memory().numpy()
    a_host[:] = env.out["mpc_actions"][:, :2].cpu().numpy()
    noise_c = np.ascontiguousarray(noise_h.transpose(0, 2, 1))   # [8][2][E]: the policy works column by column (numpy's inner
    a0_h, a1_h = a_host[:, 0], a_host[:, 1]                      # loop over a trailing dimension of 2 costs 4x more)

    def host_policy(views, i):
        ma = views["mpc_actions"]
        np.add(ma[:, 0], noise_c[i % 8, 0], out=a0_h); np.clip(a0_h, 0.0, 1.0, out=a0_h)
        np.add(ma[:, 1], noise_c[i % 8, 1], out=a1_h); np.clip(a1_h, -1.0, 1.0, out=a1_h)

    for i in range(3):
        views = env.step_host(a_host)
        host_policy(views, i)
    n_e2e = a.steps   # whole laps of the path-switch cycle, like the device-resident loop
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        views = env.step_host(a_host)
        host_policy(views, i)   # the next action is read from the host slab
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": world * E * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": E * 2 * 8, "d2h_bytes_per_step": env.slab_bytes,
           "steps": n_e2e, "ms_per_step": e2e_s / n_e2e * 1e3,
           "what": "dre_env_step_host: pinned host actions -> H2D -> robot/path half (its PT state travels during the MPC solve) -> MPC "
                   "-> one crowd-half launch whose env chunks raise flags in mapped host memory as their observations are written; "
                   "the host thread queues each chunk's D2H on a copy stream -> the whole output slab (observations, rewards, done "
                   "bits, MPC status) in pinned host memory; the next action is read from the host slab"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, nst, dt = cpu_env_steps_per_sec(a, a.cpu_envs, 12.0, threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{a.cpu_envs} of the {E} envs x {Nh} humans, {nst} steps ({dt:.1f} s), oracle/env_ref.c with OpenMP over envs",
                        "c1_single_env_1core": {"value": cpu_c1_single_env(a), "unit": UNIT,
                                                "what": "BASELINE configs[0] shape: 1 env x 6 humans, K = 4, one host core, same oracle port "
                                                        "(the reference's own Python on the API shims does ~43 env-steps/s on this shape)"}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 ORCA + f64 MPC/env", "data": "synthetic", "config": workload_config(a), "clocks": clocks,
                "e2e": e2e, "gpu_launches": launches_per_step * a.steps, "roofline": roofline, "roofline_compute": roofline_compute,
                "kernels_ms": kavg, "dedup_variant": dedup, "cpu_baseline": cpu_baseline, "wall_ms_per_step_incl_flush": t_wall / a.steps * 1e3,
                "episode_stats": {n: int(cnt[i]) for i, n in enumerate(d.env.STAT_NAMES)}, "reward_sums": [float(x) for x in sums[:3]]}
        _emit(line)
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
    else:
        run_ours(args)
