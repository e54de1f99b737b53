#!/bin/bash
OUT=gpurun_out/r2s18; mkdir -p $OUT
timeout 900 python tools/composites_timing.py > $OUT/composites_timing_r2.txt 2> $OUT/err.txt; cat $OUT/composites_timing_r2.txt; tail -2 $OUT/err.txt
