#!/bin/bash
# round 2, final tree on N = $1 GPUs exactly as the driver launches it: both bench arms under torchrun + the multi-device tests
N=${1:-2}
OUT=gpurun_out/r2finalmulti$N; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "multidevice or devices" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547"
timeout 900 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm x$N exit $?: $(tail -1 $OUT/bench_reference.json | cut -c1-220)"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default arm x$N exit $?: $(tail -1 $OUT/bench_default.json | cut -c1-260)"
tail -1 $OUT/bench_default.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('   e2e', d.get('e2e')); print('   single_call', d.get('single_call')); print('   clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))"
tail -2 $OUT/bench_default.err
