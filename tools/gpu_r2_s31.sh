#!/bin/bash
# round 2, GPU session 31: state table compacted to the TimeInterpolated components (rank mapping), full suite, timings
OUT=gpurun_out/${GB_SESSION:-r2s31}; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_stats.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
unset GB_PARITY_LOG
timeout 600 python tools/composites_timing.py > $OUT/composites_timing.txt 2>&1; cat $OUT/composites_timing.txt | tail -2
for tool in memcheck; do
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/${tool}_timeinterp.log python tools/sanitize_paths.py timeinterp > $OUT/${tool}_timeinterp.out 2>&1; echo "$tool exit $?: $(grep -h 'SUMMARY' $OUT/${tool}_timeinterp.log | tail -1)"
done
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -1 $OUT/bench_default.json | cut -c1-200
