#!/bin/bash
# tools/emu_racecheck.sh [pytest args] -- race check of the product kernels without a GPU (compute-sanitizer racecheck is closed on
# the GPU pool). The emulated build of the library (tests/host_harness/gen_emu.py) under ThreadSanitizer: every emulated CUDA thread
# is a TSan FIBER that is concurrent with all others; the only ordering TSan is told about is what CUDA's memory model gives --
# __syncthreads / __syncthreads_or / __syncwarp / cooperative-groups tile.sync() / the named barrier among their participants, kernel
# boundaries, and atomics (real __atomic operations). Shuffles and votes move values but order nothing. The CTAs of a grid are
# concurrent too (fresh stacks and shared memory per CTA). Findings are fatal (halt_on_error=1).
# Positive control: DRE_EMU_DROP_SYNCWARP=1 tools/emu_racecheck.sh -k "orca_predict and 8"   -> races reported inside
# orca_predict_kernel (the lane-group's shared scratch), because tile.sync() then waits without ordering.
# OMP_NUM_THREADS=1: the CPU oracle some tests compare with is OpenMP code TSan cannot see into. Takes ~15 minutes.
cd "$(dirname "$0")/.."
export DRE_EMU_SANITIZE=thread
export OMP_NUM_THREADS=1
export LD_PRELOAD=$(gcc -print-file-name=libtsan.so)
export TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0 history_size=4"
exec python -m pytest tests/test_emulated_library.py -q -x -s "$@"
