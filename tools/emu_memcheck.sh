#!/bin/bash
# tools/emu_memcheck.sh [pytest args] -- memcheck of EVERY product kernel and of the library's host code without a GPU: the emulated
# build of the library (tests/host_harness/gen_emu.py: the product's .cu files compiled for the CPU, CUDA threads as coroutines)
# under AddressSanitizer, driven through the real dre_* entry points by tests/test_emulated_library.py on all golden fixtures.
# "Device" allocations (cudaMalloc -> malloc), the output / state arrays the tests hand in and the dynamic shared memory of each
# launch are exact-size heap blocks, so an out-of-bounds global or shared access of a kernel is a heap-buffer-overflow report
# (positive control: an output buffer 8 floats short is reported inside orca_predict_kernel). compute-sanitizer is closed on the
# GPU pool; this is the stand-in. Summary committed as profiles/r2_emu_memcheck.txt.
cd "$(dirname "$0")/.."
# DRE_EMU_SANITIZE=undefined tools/emu_memcheck.sh runs the same suite under UndefinedBehaviorSanitizer (misaligned vector accesses,
# shifts, signed overflow, invalid enum / bool loads), findings fatal.
export DRE_EMU_SANITIZE=${DRE_EMU_SANITIZE:-address}
if [ "$DRE_EMU_SANITIZE" = "undefined" ]; then
  export LD_PRELOAD=$(gcc -print-file-name=libubsan.so)
  export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
else
  export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
  export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
fi
exec python -m pytest tests/test_emulated_library.py -q "$@"
