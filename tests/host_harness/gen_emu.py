"""TEST INFRASTRUCTURE: builds tests/host_harness/libdre_emu.so -- the WHOLE product library (every translation unit of
dr-mpc_b200/csrc: the C ABI, the host pipelines and every CUDA kernel) compiled for the CPU, so that the `-m "not gpu"` suite can
call the real `dre_*` entry points and execute the real kernel sources without a GPU.

How: each source file is copied to tests/host_harness/_gen_emu/ with a few MECHANICAL rewrites (listed below, each asserted to
match), then compiled with g++ against
  * simt_emu.h    -- the SIMT emulator: a CTA's threads are coroutines, warp / CTA collectives are barrier exchanges, the CTAs of
                     a grid run one after the other, a launch returns when the kernel has finished;
  * fake_cudart.h -- the CUDA runtime calls the host code makes (device memory = host memory, streams / events are no-ops).
Rewrites:
  1. `kernel<<<grid, block, smem, stream>>>(args);`        -> `emu::Launcher(grid, block, smem, stream)([&] { kernel(args); });`
  2. `extern __shared__ [__align__(n)] T name[];`          -> `T *name = reinterpret_cast<T *>(emu::shared_base());`
  3. `#include <cooperative_groups.h>` dropped (the two headers above are force-included in front of every unit: g++ -include)
  4. `#if defined(__CUDACC__)` / `__CUDA_ARCH__`           -> the CUDA side of the guard (`#if 1`)
  5. the two inline-PTX statements of mpc_coop.cuh         -> dre_emul_rcp_approx / dre_emul_bar_sync (simt_emu.h)
  6. `"../../include/dre.h"`                               -> the same header, one directory further up
`__shared__` variables become function-level statics (one CTA runs at a time). Nothing here is shipped; the product path is the
CUDA build and fails loudly without it.
"""
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(REPO, "dr-mpc_b200", "csrc")
HH = os.path.join(REPO, "tests", "host_harness")
GEN = os.path.join(HH, "_gen_emu")
OUT = os.path.join(HH, "libdre_emu.so")
UNITS = ["orca_kernels.cu", "mpc_kernels.cu", "env_kernels.cu", "capi.cu", "env_capi.cu", "replay_capi.cu"]


def _matching(src, i, open_ch, close_ch):
    """index of the bracket that closes src[i] (== open_ch)"""
    depth = 0
    for j in range(i, len(src)):
        if src[j] == open_ch:
            depth += 1
        elif src[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


def rewrite_launches(src):
    out, pos, n = [], 0, 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        # kernel expression: identifier, optionally followed by balanced template arguments, right before <<<
        k = i
        while k > 0 and src[k - 1].isspace():
            k -= 1
        if src[k - 1] == ">":
            depth, k2 = 0, k - 1
            while True:
                if src[k2] == ">":
                    depth += 1
                elif src[k2] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k2 -= 1
            k = k2
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kern = src[k:i].strip()
        j = src.index(">>>", i)
        cfg = src[i + 3:j]
        a0 = src.index("(", j)
        assert src[j + 3:a0].strip() == "", src[j:a0 + 20]
        a1 = _matching(src, a0, "(", ")")
        semi = src.index(";", a1)
        assert src[a1 + 1:semi].strip() == ""
        out.append(src[pos:k])
        out.append(f"emu::Launcher({cfg})([&] {{ {kern}({src[a0 + 1:a1]}); }});")
        pos = semi + 1
        n += 1
    out.append(src[pos:])
    return "".join(out), n


def transform(name, src):
    notes = {}
    src, notes["launches"] = rewrite_launches(src)
    src, notes["dyn_smem"] = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                                     lambda m: f"{m.group(1)} *{m.group(2)} = reinterpret_cast<{m.group(1)} *>(emu::shared_base());", src)
    src, notes["includes"] = re.subn(r"#include <cooperative_groups\.h>\n", "", src)
    src, notes["guards"] = re.subn(r"#if defined\(__CUDACC__\)|#if defined\(__CUDA_ARCH__\)", "#if 1", src)
    for pat, rep in (('asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));', "r = dre_emul_rcp_approx(x);"),
                     ('asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");', "dre_emul_bar_sync(bar_id, 64);")):
        if name == "mpc_coop.cuh":
            assert src.count(pat) == 1, pat
        src = src.replace(pat, rep)
    src = src.replace('"../../include/dre.h"', '"../../../include/dre.h"')
    code = "\n".join(l.split("//")[0] for l in src.splitlines())
    assert "<<<" not in code and not re.search(r"\basm\b", code), name
    return src, notes


EXPECT = {"env_kernels.cu": {"launches": 9, "dyn_smem": 1}, "orca_kernels.cu": {"dyn_smem": 3}, "mpc_kernels.cu": {"dyn_smem": 1},
          "orca_device.cuh": {"includes": 1}}


def generate():
    os.makedirs(GEN, exist_ok=True)
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh")):
            continue
        src, notes = transform(f, open(os.path.join(CSRC, f)).read())
        for k, v in EXPECT.get(f, {}).items():
            assert notes[k] == v, (f, k, notes[k], v)       # the product source changed shape: revisit the rewrite
        with open(os.path.join(GEN, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as o:
            o.write(f"// GENERATED by tests/host_harness/gen_emu.py from dr-mpc_b200/csrc/{f} -- do not edit, not committed\n" + src)


def build(force=False, sanitize=None):
    """sanitize="address": the same build under AddressSanitizer (tools/emu_memcheck.sh) -- a memcheck of every kernel and of the host
    code on every fixture: "device" allocations and the dynamic shared memory of each launch are exact-size heap blocks."""
    global OUT
    if sanitize:
        OUT = os.path.join(HH, f"libdre_emu_{sanitize}.so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HH, f) for f in ("simt_emu.h", "fake_cudart.h", "gen_emu.py")] + \
           [os.path.join(REPO, "include", "dre.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    generate()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # __shared__ -> static: redefined on the command line AFTER the CUDA headers would define it to nothing is not possible, so
    # simt_emu.h does it; -ffp-contract=off = the CUDA units' -fmad=false (x86-64 has no contraction anyway)
    flags = ["-O2", "-ffp-contract=off", "-fPIC", "-std=c++17", "-Wno-unknown-pragmas", "-Wno-attributes", "-I/usr/local/cuda/include",
             "-include", os.path.join(HH, "simt_emu.h"), "-include", os.path.join(HH, "fake_cudart.h")]
    link = []
    if sanitize:
        flags = ["-O1" if f == "-O2" else f for f in flags] + ["-g", "-fno-omit-frame-pointer", f"-fsanitize={sanitize}"]
        link = [f"-fsanitize={sanitize}"]
    procs = []
    for u in UNITS:
        cpp = os.path.join(GEN, u.replace(".cu", ".cpp"))
        obj = cpp.replace(".cpp", f".{sanitize or 'plain'}.o")
        procs.append((u, obj, subprocess.Popen([cxx] + flags + ["-c", cpp, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for u, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emulated build of {u} failed:\n{out[-6000:]}")
        objs.append(obj)
    subprocess.check_call([cxx, "-shared", "-o", OUT] + objs + link + ["-ldl"])
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, sanitize=sys.argv[1] if len(sys.argv) > 1 else None))
