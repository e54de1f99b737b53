#!/bin/bash
# round 2, GPU session 24: the two sanitizer families whose driver script had errors in session 23, + initcheck on the kernels' own scratch
OUT=gpurun_out/r2s24; mkdir -p $OUT
timeout 300 python tools/sanitize_paths.py eval timeinterp > $OUT/plain.log 2>&1; echo "plain exit $?"; tail -5 $OUT/plain.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for fam in eval timeinterp; do
    t0=$(date +%s)
    timeout 420 $CS --tool $tool --error-exitcode 9 --print-limit 10 --log-file $OUT/${tool}_${fam}.log python tools/sanitize_paths.py $fam > $OUT/${tool}_${fam}.out 2>&1
    rc=$?
    echo "$tool $fam: exit $rc, $(( $(date +%s) - t0 )) s, $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/${tool}_${fam}.log | tail -1)"
  done
done
for fam in fixed dop853 mockstream nbody; do
  t0=$(date +%s)
  timeout 420 $CS --tool initcheck --error-exitcode 9 --print-limit 10 --log-file $OUT/initcheck_${fam}.log python tools/sanitize_paths.py $fam > $OUT/initcheck_${fam}.out 2>&1
  echo "initcheck $fam: exit $?, $(( $(date +%s) - t0 )) s, $(grep -h 'ERROR SUMMARY' $OUT/initcheck_${fam}.log | tail -1)"
done
