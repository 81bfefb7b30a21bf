"""TEST INFRASTRUCTURE -- counts the floating-point operations of the flat oracle (rvo2_ref.c / mpc_ref.c compiled with
instrumented scalars, oracle/opcount/) on representative inputs and writes profiles/roofline.json, which bench.py reads for
its `roofline_compute` object (the 53 n / 2000 K figures there are estimates to be replaced by a count).

    python oracle/opcount/count.py

"flops" = add/sub + mul + div + sqrt + trig + log, one each; compares / min / max are listed but not counted.
ORCA: fp32, per solve, on crowds the free-running env oracle reaches at each BASELINE shape. MPC: fp64, per damped linear
solve of the LM (build + solve + trial costs + bookkeeping amortised), on warm re-solves along the four paths."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as O   # noqa: E402

NAMES = ["add", "mul", "div", "sqrt", "trig", "log", "cmp", "other"]


def main():
    subprocess.check_call(["make", "-C", HERE, "-B", "libopcount.so"], stdout=subprocess.DEVNULL)
    O.build()
    fast = O.lib()
    cnt = ctypes.CDLL(os.path.join(HERE, "libopcount.so"))

    def ops():
        a = (ctypes.c_uint64 * 8)()
        cnt.opcount_get(a)
        return np.array(list(a), dtype=np.float64)

    import importlib.util
    spec = importlib.util.spec_from_file_location("dre_paths", os.path.join(REPO, "dr-mpc_b200", "paths.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    tab, lens = P.pack_paths(P.continuous_task_paths())
    out = {"how": "oracle/opcount/count.py: the flat oracle compiled with instrumented scalars; flops = add + mul + div + sqrt + trig + log",
           "orca": {}, "mpc": {}}

    # ---------------- ORCA: states the env reaches (uncounted fast oracle), then one counted crowd pass ----------------
    for Nh, E, crh in ((6, 256, 5.0), (10, 256, 5.0), (20, 256, 5.0), (50, 64, 8.0)):
        O._lib = fast
        env = O.EnvRef(E, Nh, tab, lens, circle_radius_humans=crh)
        a = env.reset()
        for _ in range(25):
            a = env.step(a["mpc_actions"][:, :2].copy())
        O._lib = cnt
        cnt.opcount_reset()
        v, st = O.orca_crowd_pass(a["hpos"], a["hvel"], a["hgoal"], a["rpose"][:, :2], nthreads=1, want_stats=True)
        c = ops()
        n = E * Nh
        flops = c[:6].sum() / n
        out["orca"][f"n{Nh}"] = {
            "neighbours": Nh, "flops_per_solve": flops, "breakdown_per_solve": {k: c[i] / n for i, k in enumerate(NAMES)},
            "survey_formula_53n_plus_lp1": 53.0 * Nh + 10.0 * float(st["lp1_work"].mean()) + 20.0 * float(st["n_lp1"].mean()),
            "mean_lp1_calls": float(st["n_lp1"].mean()), "lp3_fraction": float((st["lp2_fail"] < st["n_lines"]).mean()),
            "sample": f"{E} envs x {Nh} humans after 25 free-running steps, circle_radius_humans {crh}"}
        print(f"ORCA n={Nh}: {flops:.0f} flops/solve (survey formula {out['orca'][f'n{Nh}']['survey_formula_53n_plus_lp1']:.0f})")

    # ---------------- MPC: iteration0 + 4 warm re-solves (BASELINE config 5), counted ----------------
    for K, N in ((4, 512), (10, 256), (20, 128), (30, 96), (40, 64)):
        rng = np.random.default_rng(K)
        node = rng.integers(0, lens[0] - 1, N)
        pose = tab[0][:, node].T + rng.uniform(-0.3, 0.3, (N, 3))
        idx = node.astype(np.float64)
        O._lib = cnt
        mb = O.MpcBatch(N, O.mpc_params(K=K), tab, lens)
        tot = np.zeros(8); solves = iters = evals = acts = 0
        first = None
        for step in range(5):
            cnt.opcount_reset()
            u, st = mb.act(np.zeros(N, np.int32), idx, pose, nthreads=1)
            c = ops()
            if step == 0:
                first = (c[:6].sum() / N, float(st["lm_solves"].mean()))
            else:
                tot += c; solves += int(st["lm_solves"].sum()); iters += int(st["iterations"].sum()); evals += int(st["cost_evals"].sum()); acts += N
            v, w = u[:, 0], u[:, 1]
            w_ = np.where(w == 0, 1e-12, w)
            dx = (v / w_) * np.sin(w_ * 0.25); dy = (v / w_) * (1 - np.cos(w_ * 0.25))
            cth, sth = np.cos(pose[:, 2]), np.sin(pose[:, 2])
            pose = pose + np.stack([cth * dx - sth * dy, sth * dx + cth * dy, w * 0.25], 1)
            idx = np.minimum(idx + v * 0.25 / 0.2, lens[0] - 1)
        out["mpc"][f"K{K}"] = {
            "horizon": K, "flops_per_damped_solve": tot[:6].sum() / solves, "flops_per_act_warm": tot[:6].sum() / acts,
            "breakdown_per_damped_solve": {k: tot[i] / solves for i, k in enumerate(NAMES)},
            "mean_lm_solves_warm": solves / acts, "mean_lm_iterations_warm": iters / acts, "mean_cost_evals_warm": evals / acts,
            "flops_per_act_iteration0": first[0], "mean_lm_solves_iteration0": first[1], "survey_estimate_2000K": 2000.0 * K,
            "sample": f"{N} problems on path 0, poses +-0.3 m / rad off the path, 4 warm re-solves after the iteration0 solve"}
        print(f"MPC K={K}: {out['mpc'][f'K{K}']['flops_per_damped_solve']:.0f} flops per damped solve (survey estimate {2000 * K}), "
              f"{solves / acts:.1f} damped solves per warm act")
    O._lib = fast
    with open(os.path.join(REPO, "profiles", "roofline.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote profiles/roofline.json")


if __name__ == "__main__":
    main()
