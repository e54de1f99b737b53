#!/bin/bash
# round 2: trajectory-output workloads through ONE call over N devices (every device reads its slice back over its own PCIe link)
N=${1:-8}
OUT=gpurun_out/r2multitraj$N; mkdir -p $OUT
free -g | head -2 > $OUT/mem.txt
for w in c1 c1x c2; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "$w x$N: $(tail -1 $OUT/bench_$w.json | cut -c1-150)"; tail -1 $OUT/bench_$w.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('   single_call', d.get('single_call'))"
  tail -1 $OUT/bench_$w.err
done
# the same three on one device of this box, for the 1-GPU point of the curve
for w in c1 c1x c2; do
  timeout 1500 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench1_$w.json 2> $OUT/bench1_$w.err
  tail -1 $OUT/bench1_$w.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$w x1 single_call', d.get('single_call'))"
done
cat $OUT/mem.txt
