#!/usr/bin/env python
"""SASS instruction count of one kernel grouped by the source FUNCTION the line belongs to (inlined code included):
python tools/sass_funcs.py <obj> <kernel-substr> [hot_limit_hex]   -- hot_limit: only count addresses below it"""
import sys, re, collections, subprocess, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
limit = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
out = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
fn_of = {}
def funcs(path):
    if path in fn_of: return fn_of[path]
    tab = []
    try:
        cur = "?"
        for i, ln in enumerate(open(path), 1):
            m = re.match(r'^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|DRE_HD|__global__|__SM_30_INTRINSICS_DECL__|static __device__)[^;(]*?([A-Za-z_0-9]+)\s*\(', ln)
            if m and not ln.rstrip().endswith(';'): cur = m.group(1)
            tab.append(cur)
    except OSError:
        pass
    fn_of[path] = tab
    return tab
cnt = collections.Counter(); cur = None; fn = None
for ln in out.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', ln)
    if m: fn = m.group(1); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(\S+)', ln)
    if fn and pat in fn and m and cur:
        if int(m.group(1), 16) >= limit: continue
        t = funcs(cur[0]); name = t[cur[1] - 1] if cur[1] - 1 < len(t) else "?"
        cnt[(os.path.basename(cur[0]), name)] += 1
tot = sum(cnt.values()); print("total", tot, "instr =", tot * 16, "bytes")
for k, v in cnt.most_common(40): print(f"{v:6d}  {k[0]}:{k[1]}")
