#!/bin/bash
# round 2, GPU session 29: + prefetch of the next step's TimeInterpolated state row
OUT=gpurun_out/r2s29; mkdir -p $OUT
timeout 600 python tools/ti_profile.py > $OUT/plain.log 2>&1; tail -3 $OUT/plain.log
timeout 900 python -m pytest tests -m gpu -q -x -k "timeinterp" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
