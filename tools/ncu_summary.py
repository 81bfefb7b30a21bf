"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/launches.csv, gpurun_out/<rep>.ncu-rep) into the
small text summaries committed under profiles/. Usage: python tools/ncu_summary.py <tag> [rep-file]"""
import collections
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else None
out_dir = os.path.join(REPO, "profiles")
os.makedirs(out_dir, exist_ok=True)

OURS = ("ha_step_kernel", "pt_step_kernel", "mpc_coop_kernel", "mpc_act_kernel", "mpc_stats_kernel", "orca_predict_kernel",
        "env_reset_kernel", "fma_chain_kernel", "mpc_reset_kernel")


def short(name):
    for o in OURS:
        if o in name:
            i = name.index(o)
            return name[i:].split("(")[0]
    return name.split("(")[0][-70:]


launches = os.path.join(REPO, "gpurun_out", "launches.csv")
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    step = {k: v for k, v in agg.items() if any(o in k for o in ("ha_step", "pt_step", "mpc_coop", "mpc_act", "mpc_stats"))}
    tot = sum(v[1] for v in step.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dedup)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':50s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share_of_step_kernels':>22s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sh = f"{v[1] / tot * 100:6.1f}%" if k in step else "   (not on the step path)"
            f.write(f"{k:50s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.2f} {sh:>22s}\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.txt")).read())

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
    seen = {}
    traffic = {}
    issue = {}
    with open(os.path.join(out_dir, f"{tag}_kernels.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dedup)\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            if name in seen:
                continue
            seen[name] = 1
            f.write(f"\n== {name} ==\n")
            for w in want:
                if w in idx:
                    f.write(f"  {w:75s} {r[idx[w]]:>18s} {units[idx[w]]}\n")
            st = [(h, float(r[i].replace(",", ""))) for h, i in idx.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
            st.sort(key=lambda x: -x[1])
            f.write("  top stall reasons (warps stalled per issue-active cycle):\n")
            for h, v in st[:7]:
                f.write(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}\n")
            def num(key):
                v = float(r[idx[key]].replace(",", ""))
                u = units[idx[key]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            key = "ha_step" if "ha_step" in name else ("mpc_solve" if "mpc_coop" in name or "mpc_act" in name else ("pt_step" if "pt_step" in name else name))
            traffic[key] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            issue[key] = {"issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
                          "warps_active_pct": float(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
                          "threads_per_warp_inst": float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]].replace(",", "")),
                          "warp_instructions": float(r[idx["smsp__inst_executed.sum"]].replace(",", "")), "source": f"profiles/{tag}_kernels.txt"}
    json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
    json.dump(issue, open(os.path.join(out_dir, "issue.json"), "w"), indent=1)
    print(open(os.path.join(out_dir, f"{tag}_kernels.txt")).read()[:3000])
    print(traffic)
