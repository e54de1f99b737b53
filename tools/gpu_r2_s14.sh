#!/bin/bash
OUT=gpurun_out/r2s14; mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
export GALA_B200_LIB=$PWD/gala_b200/libgala_b200_sync.so
for k in 1 2 4 8; do
  GB_D8_BLOCKSYNC=$k GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_sync$k.json 2> $OUT/c2_sync$k.err; echo "sync every $k: $(tail -1 $OUT/c2_sync$k.err)"
done
GB_D8_BLOCKSYNC=4 timeout 600 python -m pytest tests -m gpu -q -k "dop853" 2>&1 | tail -1
