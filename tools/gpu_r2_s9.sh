#!/bin/bash
# round 2, GPU session 9: what the driver runs at round end, on the final build: the whole -m gpu suite, smoke(), the default bench
OUT=gpurun_out/r2s9; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_distributions.txt
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
unset GB_PARITY_LOG
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2> $OUT/bench_default.time; tail -1 $OUT/bench_default.json | cut -c1-300; tail -3 $OUT/bench_default.time
