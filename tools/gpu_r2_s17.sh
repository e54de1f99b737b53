#!/bin/bash
OUT=gpurun_out/r2s17; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for w in c3sg c3sgd; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $OUT/bench_r2_$w.json 2> $OUT/bench_r2_$w.err; echo "$w: $(tail -1 $OUT/bench_r2_$w.json | cut -c1-150)"
done
