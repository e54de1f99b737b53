#!/bin/bash
# round 2, GPU session 23: compute-sanitizer (memcheck, racecheck, synccheck) over one tiny call of every kernel family
OUT=gpurun_out/r2s23; mkdir -p $OUT
timeout 300 python tools/sanitize_paths.py > $OUT/plain.log 2>&1; echo "plain exit $?"; tail -12 $OUT/plain.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for fam in eval fixed dop853 extrema timeinterp nbody mockstream lyapunov; do
    t0=$(date +%s)
    timeout 420 $CS --tool $tool --error-exitcode 9 --print-limit 10 --log-file $OUT/${tool}_${fam}.log python tools/sanitize_paths.py $fam > $OUT/${tool}_${fam}.out 2>&1
    rc=$?
    echo "$tool $fam: exit $rc, $(( $(date +%s) - t0 )) s, $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/${tool}_${fam}.log | tail -1)"
  done
done
