#!/bin/bash
OUT=gpurun_out/r2s19; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 900 python tools/composites_timing.py > $OUT/composites_timing_r2.txt 2> $OUT/err.txt; cat $OUT/composites_timing_r2.txt; tail -2 $OUT/err.txt
