#!/usr/bin/env python
"""SASS instruction count per source line of one kernel: python tools/sass_lines.py <obj> <kernel-substr> [top]"""
import sys, re, collections, subprocess, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
out = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; fn = None
for ln in out.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', ln)
    if m: fn = m.group(1); continue
    if fn and pat in fn and re.match(r'\s+/\*[0-9a-f]{4,5}\*/', ln): cnt[cur] += 1
tot = sum(cnt.values()); print('total instr', tot, '=', tot * 16, 'bytes')
perfile = collections.Counter()
for (f, l), v in cnt.items(): perfile[f] += v
print(dict(perfile))
for k, v in cnt.most_common(top): print(k, v)
