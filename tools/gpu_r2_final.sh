#!/bin/bash
# round 2, final bench lines: every workload with >= 20 steps, the e2e leg, the single-call leg and the CPU baseline
# (1 thread and all cores), one JSON line each -> profiles/bench_r2_<workload>.json
OUT=gpurun_out/r2final; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
for w in ${WORKLOADS:-headline c1 c1x c2 c4 c5 c3 c3d c3sg c3sgd}; do
  timeout 1500 python bench.py --workload $w --steps 20 --warmup 5 > $OUT/bench_r2_$w.json 2> $OUT/bench_r2_$w.err
  echo "$w: $(tail -1 $OUT/bench_r2_$w.json | cut -c1-160)"; tail -1 $OUT/bench_r2_$w.err
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_r2_reference_arm_headline.json 2> $OUT/ref.err; tail -1 $OUT/bench_r2_reference_arm_headline.json | cut -c1-200
