#!/bin/bash
OUT=gpurun_out/r2s15; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multidevice.py -m gpu -q -k "concurrent or validation or fewer" > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
