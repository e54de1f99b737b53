#!/bin/bash
# round 2, GPU session 4: C2 with k6..k9 parked in shared memory, mock-stream pipeline after the reorder
OUT=gpurun_out/r2s4; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_stats.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "step_statistics or pipelined or mockstream or one_burst or c2_dop853 or dop853" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log; cat $OUT/parity_stats.txt | head -5
unset GB_PARITY_LOG
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_timing.json 2> $OUT/c2_timing.err; echo "c2 timing: $(tail -1 $OUT/c2_timing.json | cut -c1-140)"; tail -2 $OUT/c2_timing.err
timeout 600 python bench.py --workload c2 $B > $OUT/c2.json 2> $OUT/c2.err; echo "c2: $(tail -1 $OUT/c2.json | cut -c1-140)"
GB_STREAM_TRACE=1 timeout 600 python bench.py --workload c3 $B > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3: $(tail -1 $OUT/bench_c3.json | cut -c1-140)"; tail -2 $OUT/bench_c3.err
GB_STREAM_TRACE=1 timeout 600 python bench.py --workload c3d $B > $OUT/bench_c3d.json 2> $OUT/bench_c3d.err; echo "c3d: $(tail -1 $OUT/bench_c3d.json | cut -c1-140)"; tail -1 $OUT/bench_c3d.err
