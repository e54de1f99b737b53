#!/bin/bash
# round 2, GPU session 1: parity after the C-ABI refactor, DOP853 dense-kernel A/B (occupancy variants), ncu of the
# new dense kernel with per-instruction source counters, save_all roofline workloads.
OUT=gpurun_out/r2s1; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
export GB_PARITY_LOG=$PWD/$OUT/parity_distributions.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
unset GB_PARITY_LOG
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
for v in default minb2 minb4; do
  [ $v = default ] && unset GALA_B200_LIB || export GALA_B200_LIB=$PWD/gala_b200/libgala_b200_$v.so
  timeout 600 python bench.py --workload c2 $B > $OUT/c2_$v.json 2> $OUT/c2_$v.err; echo "c2 $v: $(tail -1 $OUT/c2_$v.json | cut -c1-140)"
done
unset GALA_B200_LIB
for w in headline c1 c1x c4; do
  timeout 600 python bench.py --workload $w $B > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w: $(tail -1 $OUT/bench_$w.json | cut -c1-140)"; tail -2 $OUT/bench_$w.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dop853 -s 3 -c 1 -o $OUT/prof_dop853 -f python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/ncu_d8.log 2>&1
python tools/ncu_summary.py $OUT/prof_dop853.ncu-rep "ncu --set full --clock-control none, k_dop853_dyn<MW2022, static, dense> warp-cooperative dense output, 303,104 orbits x 1000 output times (r2s1)" > $OUT/ncu_r2_dop853.txt 2> $OUT/ncu_summ.err
ncu -i $OUT/prof_dop853.ncu-rep --page source --csv > $OUT/prof_dop853_source.csv 2> $OUT/ncu_src.err
gzip -f $OUT/prof_dop853_source.csv; ls -la $OUT/prof_dop853.ncu-rep $OUT/prof_dop853_source.csv.gz
rm -f $OUT/prof_dop853.ncu-rep
grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|warps_active|dram__bytes|thread_inst_executed_per" $OUT/ncu_r2_dop853.txt
