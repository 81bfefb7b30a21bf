#!/bin/bash
# tools/run_configs.sh N OUT [STEPS] -- every BASELINE.json config through bench.py on N GPUs of this box, one JSON line each,
# appended to OUT (profiles/configs_r2.jsonl is the committed copy). Each line carries `parity` (the config's own oracle gate).
N=${1:-1}; OUT=${2:-gpurun_out/configs.jsonl}; STEPS=${3:-100}
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" >> "$OUT" 2>> "$OUT.err"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus "$N" "$@" >> "$OUT" 2>> "$OUT.err"; fi
  echo "rc=$? $*" >> "$OUT.err"
}
[ "$N" = "1" ] && run --config C2 --steps 1000 --warmup 50
run --config C3 --scaling weak --steps "$STEPS" --warmup 20
[ "$N" != "1" ] && run --config C3 --scaling strong --steps "$STEPS" --warmup 20 --no-dedup
run --config C4 --scaling strong --steps "$STEPS" --warmup 10
run --config C4 --scaling strong --steps "$STEPS" --warmup 10 --neighbor-dist 3
for K in 10 20 30 40; do run --config C5 --scaling strong --horizon $K --steps 4 --warmup 3; done
