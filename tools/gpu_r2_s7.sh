#!/bin/bash
# round 2, GPU session 7: dense DOP853 writing the caller's layout directly (no scratch, no transpose) vs the scratch + transpose form
OUT=gpurun_out/r2s7; mkdir -p $OUT
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_scratch.json 2> $OUT/c2_scratch.err; echo "c2 scratch+transpose: $(tail -1 $OUT/c2_scratch.json | cut -c1-140)"; tail -1 $OUT/c2_scratch.err
GB_D8_DIRECT=1 GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_direct.json 2> $OUT/c2_direct.err; echo "c2 direct: $(tail -1 $OUT/c2_direct.json | cut -c1-140)"; tail -1 $OUT/c2_direct.err
GB_D8_DIRECT=1 GB_D8_NOSORT=1 GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_direct_nosort.json 2> $OUT/c2_direct_nosort.err; echo "c2 direct nosort: $(tail -1 $OUT/c2_direct_nosort.json | cut -c1-140)"; tail -1 $OUT/c2_direct_nosort.err
GB_D8_DIRECT=1 timeout 900 python -m pytest tests -m gpu -q -k "dop853 or multidevice or extrema or smoke" > $OUT/pytest_direct.log 2>&1; tail -3 $OUT/pytest_direct.log
GB_D8_DIRECT=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:k_dop853 -s 3 -c 1 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call 2>&1 | grep -E "dram__|gpu__time|lts__" 
