#!/bin/bash
# round 2, GPU session 5: the whole parity suite on the current build
OUT=gpurun_out/r2s5; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_distributions.txt
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -8 $OUT/pytest.log
grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
