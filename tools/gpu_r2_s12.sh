#!/bin/bash
OUT=gpurun_out/r2s12; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_extrema.py tests/test_gpu_multidevice.py tests/test_gpu_known_answers.py -m gpu -q > $OUT/pytest.log 2>&1; tail -12 $OUT/pytest.log
