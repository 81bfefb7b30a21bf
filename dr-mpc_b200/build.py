"""Builds dr-mpc_b200/libdre.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

ORCA / env translation units are compiled with -fmad=false so that every fp32/fp64 operation rounds
once, like the SSE builds of RVO2 / numpy the reference runs on (bit-exact ORCA parity); the MPC
translation unit keeps FMA contraction (fp64 throughput; parity tolerance 1e-6 on controls).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdre.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]

# (source, extra flags)
UNITS = [
    ("orca_kernels.cu", ["-fmad=false"]),
    ("mpc_kernels.cu", []),
    ("env_kernels.cu", ["-fmad=false"]),
    ("capi.cu", []),
    ("env_capi.cu", []),
    ("replay_capi.cu", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libdre.so cannot be built")


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dre.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, logs = [], []
    procs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, "-ccbin", _host_cxx()] + ARCH + COMMON + extra + ["-c", path, "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        logs.append("$ " + " ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(logs[-1])
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed (see above)")
    link = [nvcc, "-ccbin", _host_cxx()] + ARCH + ["-shared", "-o", OUT] + objs
    subprocess.check_call(link)
    if verbose:
        print("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
