#!/bin/bash
# round 2, GPU session 30: full GPU suite + composite timings + default bench on the tree with the TimeInterpolated state table
OUT=gpurun_out/r2s30; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_stats.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
unset GB_PARITY_LOG
timeout 600 python tools/composites_timing.py > $OUT/composites_timing.txt 2>&1; cat $OUT/composites_timing.txt | tail -3
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -1 $OUT/bench_default.json | cut -c1-300
