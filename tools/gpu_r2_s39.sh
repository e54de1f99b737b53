#!/bin/bash
# round 2, GPU session 39: sanitizer over the time-dependent paths added late in the round (state table, Hessian, mock streams)
OUT=gpurun_out/r2s39; mkdir -p $OUT
timeout 300 python tools/sanitize_paths.py timeinterp > $OUT/plain.log 2>&1; echo "plain exit $?"; tail -2 $OUT/plain.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  t0=$(date +%s)
  timeout 420 $CS --tool $tool --error-exitcode 9 --print-limit 10 --log-file $OUT/${tool}_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/${tool}_timeinterp.out 2>&1
  echo "$tool timeinterp: exit $?, $(( $(date +%s) - t0 )) s, $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/${tool}_timeinterp.log | tail -1)"
done
