#!/bin/bash
# round 2, GPU session 26: per-step TimeInterpolated state shared by the CTA in leapfrog / Ruth4
OUT=gpurun_out/r2s26; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "timeinterp or parity or known or multidevice or point_quantities" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
timeout 600 python tools/composites_timing.py > $OUT/composites_timing.txt 2>&1; cat $OUT/composites_timing.txt | tail -8
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/racecheck_timeinterp.out 2>&1; echo "racecheck exit $?: $(grep -h 'RACECHECK SUMMARY' $OUT/racecheck_timeinterp.log | tail -1)"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $OUT/synccheck_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/synccheck_timeinterp.out 2>&1; echo "synccheck exit $?: $(grep -h 'ERROR SUMMARY' $OUT/synccheck_timeinterp.log | tail -1)"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/memcheck_timeinterp.out 2>&1; echo "memcheck exit $?: $(grep -h 'ERROR SUMMARY' $OUT/memcheck_timeinterp.log | tail -1)"
