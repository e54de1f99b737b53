#!/bin/bash
# round 2, GPU session 3: in-place timing of the C2 call, CTA-size variants, mock-stream pipeline trace
OUT=gpurun_out/r2s3; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_stats.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "step_statistics or extrema or pipelined or multidevice or origin" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log; cat $OUT/parity_stats.txt
unset GB_PARITY_LOG
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_timing.json 2> $OUT/c2_timing.err; echo "c2 timing: $(tail -1 $OUT/c2_timing.json | cut -c1-140)"; tail -4 $OUT/c2_timing.err
for v in "b384 384" "b192 192"; do set -- $v
  GALA_B200_LIB=$PWD/gala_b200/libgala_b200_$1.so GB_D8_BLOCK=$2 GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_$1.json 2> $OUT/c2_$1.err; echo "c2 $1: $(tail -1 $OUT/c2_$1.json | cut -c1-140)"; tail -2 $OUT/c2_$1.err
done
GB_STREAM_TRACE=1 timeout 600 python bench.py --workload c3 $B > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3: $(tail -1 $OUT/bench_c3.json | cut -c1-140)"; tail -3 $OUT/bench_c3.err
GB_STREAM_TRACE=1 GB_STREAM_CHUNKS=4 timeout 600 python bench.py --workload c3 $B > $OUT/bench_c3_4.json 2> $OUT/bench_c3_4.err; echo "c3 4 chunks: $(tail -1 $OUT/bench_c3_4.json | cut -c1-140)"; tail -1 $OUT/bench_c3_4.err
GB_STREAM_TRACE=1 timeout 600 python bench.py --workload c3d $B > $OUT/bench_c3d.json 2> $OUT/bench_c3d.err; echo "c3d: $(tail -1 $OUT/bench_c3d.json | cut -c1-140)"; tail -1 $OUT/bench_c3d.err
