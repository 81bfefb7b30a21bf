#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump per source line: stall samples, warp instructions, thread efficiency.
usage: ncu -i rep --page source --csv --kernel-name regex:X > src.csv ; python tools/src_hot.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
stall_tot = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    def num(k):
        try: return float(d.get(k, "0").replace(",", ""))
        except ValueError: return 0.0
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, {"src": r[1].strip()[:90], "samples": 0, "inst": 0, "tinst": 0, "stalls": collections.Counter()})
    a["samples"] += num("# Samples"); a["inst"] += num("Instructions Executed"); a["tinst"] += num("Thread Instructions Executed")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = num(k); a["stalls"][k[6:]] += v; stall_tot[k[6:]] += v
S = sum(a["samples"] for a in agg.values()); I = sum(a["inst"] for a in agg.values())
print(f"total samples {S:.0f}  warp-instructions {I:.0f}")
print("stall totals:", ", ".join(f"{k} {v / max(S,1) * 100:.1f}%" for k, v in stall_tot.most_common(9)))
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'thr/inst':>8s}  top stalls | source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = " ".join(f"{k}:{v / max(a['samples'],1) * 100:.0f}" for k, v in a["stalls"].most_common(3))
    print(f"{f + ':' + str(ln):28s} {a['samples'] / max(S,1) * 100:6.2f} {a['inst'] / max(I,1) * 100:6.2f} {a['tinst'] / max(a['inst'],1):8.1f}  {st:40s} | {a['src']}")
# per-file totals
pf = collections.Counter(); pi = collections.Counter()
for (f, ln), a in agg.items(): pf[f] += a["samples"]; pi[f] += a["inst"]
print("per file:", ", ".join(f"{f} samp {pf[f] / max(S,1) * 100:.1f}% inst {pi[f] / max(I,1) * 100:.1f}%" for f in pf))
