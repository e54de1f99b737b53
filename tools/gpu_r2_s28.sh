#!/bin/bash
# round 2, GPU session 28: TimeInterpolated state tabulated per step by a pre-pass kernel (leapfrog / Ruth4)
OUT=gpurun_out/r2s28; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "timeinterp or parity or known or multidevice or point_quantities" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
timeout 600 python tools/composites_timing.py > $OUT/composites_timing.txt 2>&1; cat $OUT/composites_timing.txt | tail -8
for tool in memcheck racecheck; do
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/${tool}_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/${tool}_timeinterp.out 2>&1; echo "$tool exit $?: $(grep -h 'SUMMARY' $OUT/${tool}_timeinterp.log | tail -1)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 2 -c 1 -o $OUT/prof_ti -f python tools/ti_profile.py > $OUT/ncu_ti.log 2>&1
python tools/ncu_summary.py $OUT/prof_ti.ncu-rep "k_leapfrog<SIG_GENERIC_TI>, NFW + TimeInterpolated Plummer (moving origin), 3,031,040 orbits x 1000 steps, state table (r2s28)" > $OUT/ncu_r2_leapfrog_ti.txt 2> $OUT/summ.err
ncu -i $OUT/prof_ti.ncu-rep --page source --csv > $OUT/prof_ti_source.csv 2>/dev/null; gzip -f $OUT/prof_ti_source.csv
rm -f $OUT/prof_ti.ncu-rep
grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|thread_inst_executed_per|warps_active|issue_active" $OUT/ncu_r2_leapfrog_ti.txt | awk '{print "   ", $1, $NF}'
