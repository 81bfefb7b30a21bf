"""tools/ptxas_summary.py [log] -> one line per kernel of the build (dr-mpc_b200/build/ptxas.log, written by build.py with
-Xptxas -v): registers, stack frame, spills, barriers, static shared / constant memory. No GPU needed."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "dr-mpc_b200", "build", "ptxas.log")
rows, cur = [], None
for line in open(log):
    m = re.search(r"Compiling entry function '([^']+)' for 'sm_100a'", line)
    if m:
        cur = {"name": m.group(1), "regs": None, "stack": 0, "sst": 0, "sld": 0, "bar": 0, "smem": 0, "cmem": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur["regs"] is None:                       # the entry function's own line precedes its "Used ..." line
        cur["stack"], cur["sst"], cur["sld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?", line)
    if m and cur["regs"] is None:
        cur["regs"], cur["bar"] = int(m.group(1)), int(m.group(2) or 0)
        s = re.search(r"(\d+) bytes smem", line)
        c = re.findall(r"(\d+) bytes cmem\[\d+\]", line)
        cur["smem"] = int(s.group(1)) if s else 0
        cur["cmem"] = sum(int(x) for x in c)
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"{'registers':>9} {'stack B':>8} {'spill st/ld B':>14} {'barriers':>8} {'smem B':>7}  kernel")
for r, n in sorted(zip(rows, names), key=lambda x: x[1]):
    n = re.sub(r"\((?!anonymous namespace\)).*", "", n).replace("void ", "")
    print(f"{r['regs']:>9} {r['stack']:>8} {str(r['sst']) + ' / ' + str(r['sld']):>14} {r['bar']:>8} {r['smem']:>7}  {n}")
