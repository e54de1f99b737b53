#!/bin/bash
# round 2, GPU session 21: ncu --set full of the FINAL dense DOP853 kernel (k6..k9 parked in shared memory) + launch list of a C2 run
OUT=gpurun_out/r2s21; mkdir -p $OUT
FP64M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"
timeout 900 ncu --set full --metrics $FP64M --clock-control none --import-source on -k regex:k_dop853 -s 3 -c 1 -o $OUT/prof_dop853 -f python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/ncu_d8.log 2>&1
python tools/ncu_summary.py $OUT/prof_dop853.ncu-rep "ncu --set full --clock-control none, k_dop853_dyn<MW2022, static, dense>, FINAL kernel of round 2, 303,104 orbits x 1000 output times (r2s21)" > $OUT/ncu_r2_dop853_final.txt 2> $OUT/summ.err
ncu -i $OUT/prof_dop853.ncu-rep --page source --csv > $OUT/prof_dop853_source.csv 2>/dev/null; gzip -f $OUT/prof_dop853_source.csv; rm -f $OUT/prof_dop853.ncu-rep
grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|warps_active|dram__bytes|thread_inst_executed_per|local_op" $OUT/ncu_r2_dop853_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/launches_c2.log 2>&1
