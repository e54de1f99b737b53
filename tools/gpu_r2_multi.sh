#!/bin/bash
# round 2, multi-GPU session (N = $1 GPUs): one C-ABI call sharded over the devices (bit identity with one device),
# weak-scaling bench under torchrun + the strong-scaling single-call measurement on rank 0.
N=${1:-2}
OUT=gpurun_out/r2multi$N; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multidevice.py -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
for w in headline c5; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $w --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "$w x$N: $(tail -1 $OUT/bench_$w.json | cut -c1-200)"; tail -1 $OUT/bench_$w.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('   e2e', d.get('e2e')); print('   single_call', d.get('single_call'))"
  tail -2 $OUT/bench_$w.err
done
