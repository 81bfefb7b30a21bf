"""tools/sass_mix.py <object> <mangled kernel> [<mangled callee> ...] -> static SASS instruction mix of a kernel (and of the
non-inlined device functions it calls, when named): instruction count per class. No GPU needed (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys

CLASSES = [("fp32 arithmetic", r"^(FADD|FMUL|FFMA|FMNMX|FSEL|FSET|FSETP|FCHK|MUFU|FSWZ)"), ("fp64 arithmetic", r"^(DADD|DMUL|DFMA|DSETP|DMNMX)"),
           ("integer / logic / moves", r"^(IADD|IADD3|IMAD|ISETP|LOP3|LOP|SHF|LEA|MOV|SEL|PRMT|IMNMX|POPC|FLO|BREV|I2F|F2I|F2F|I2I|IABS|VIADD|VIMNMX|UMOV|UIADD3|ULOP3|USHF|UISETP|ULEA|UIMAD|USEL|UPOPC|UFLO|UPRMT|PLOP3|P2R|R2P|CS2R|S2R|S2UR|R2UR|UPLOP3|LDC|LDCU|ULDC|VOTEU|UMOV)"),
           ("shared memory", r"^(LDS|STS|ATOMS|LDSM)"), ("global / local memory", r"^(LDG|STG|LD|ST|LDL|STL|ATOMG|ATOM|RED|CCTL|MEMBAR|ERRBAR)"),
           ("warp shuffles / votes / redux", r"^(SHFL|VOTE|REDUX|MATCH)"), ("barriers / warp sync", r"^(BAR|WARPSYNC|NANOSLEEP|DEPBAR|ENDCOLLECTIVE)"),
           ("control flow", r"^(BRA|BRX|JMP|CALL|RET|EXIT|BSSY|BSYNC|BREAK|YIELD|NOP|BMOV|BPT|KILL)")]
obj, funs = sys.argv[1], sys.argv[2:]
total = collections.Counter()
ops = collections.Counter()
for f in funs:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", f, obj], capture_output=True, text=True).stdout
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1).split(".")[0]
        ops[op] += 1
        for name, pat in CLASSES:
            if re.fullmatch(pat[1:], op):
                total[name] += 1
                break
        else:
            total["other"] += 1
n = sum(total.values())
print(f"{n} SASS instructions")
for name, c in total.most_common():
    print(f"  {c:6d}  {100.0 * c / n:5.1f} %  {name}")
print("  top opcodes:", ", ".join(f"{o} {c}" for o, c in ops.most_common(14)))
