#!/bin/bash
# round 2, GPU session 27: where the TimeInterpolated leapfrog kernel spends its time (ncu source page)
OUT=gpurun_out/r2s27; mkdir -p $OUT
timeout 300 python tools/ti_profile.py > $OUT/plain.log 2>&1; tail -3 $OUT/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 2 -c 1 -o $OUT/prof_ti -f python tools/ti_profile.py > $OUT/ncu_ti.log 2>&1
python tools/ncu_summary.py $OUT/prof_ti.ncu-rep "k_leapfrog<SIG_GENERIC_TI>, NFW + TimeInterpolated Plummer, 3,031,040 orbits x 1000 steps (r2s27)" > $OUT/ncu_r2_leapfrog_ti.txt 2> $OUT/summ.err
ncu -i $OUT/prof_ti.ncu-rep --page source --csv > $OUT/prof_ti_source.csv 2>/dev/null; gzip -f $OUT/prof_ti_source.csv
rm -f $OUT/prof_ti.ncu-rep
grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|dram__bytes|thread_inst_executed_per|warps_active|issue_active" $OUT/ncu_r2_leapfrog_ti.txt | awk '{print "   ", $1, $NF}'
