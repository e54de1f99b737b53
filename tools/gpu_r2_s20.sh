#!/bin/bash
OUT=gpurun_out/r2s20; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
