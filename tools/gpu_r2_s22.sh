#!/bin/bash
# round 2, GPU session 22: dense DOP853 with the caller's time grid copied to shared memory (A/B against global)
OUT=gpurun_out/r2s22; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_stats.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "dop853 or mockstream or step_statistics" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
unset GB_PARITY_LOG
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
for rep in 1 2; do
GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_smem_$rep.json 2> $OUT/c2_smem_$rep.err; echo "c2 tgrid smem: $(tail -1 $OUT/c2_smem_$rep.json | cut -c1-140)"; tail -2 $OUT/c2_smem_$rep.err
GB_D8_TGRID_GLOBAL=1 GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_glob_$rep.json 2> $OUT/c2_glob_$rep.err; echo "c2 tgrid global: $(tail -1 $OUT/c2_glob_$rep.json | cut -c1-140)"; tail -2 $OUT/c2_glob_$rep.err
done
timeout 600 python bench.py --workload c3d $B > $OUT/bench_c3d.json 2> $OUT/bench_c3d.err; echo "c3d: $(tail -1 $OUT/bench_c3d.json | cut -c1-140)"
