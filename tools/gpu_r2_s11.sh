#!/bin/bash
# round 2, GPU session 11: mock-stream bench lines again (traffic now recorded), new multi-device edge test
OUT=gpurun_out/r2s11; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multidevice.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for w in c3 c3d c3sg c3sgd; do
  timeout 1500 python bench.py --workload $w --steps 20 --warmup 5 > $OUT/bench_r2_$w.json 2> $OUT/bench_r2_$w.err
  echo "$w: $(tail -1 $OUT/bench_r2_$w.json | cut -c1-160)"; tail -1 $OUT/bench_r2_$w.err
done
