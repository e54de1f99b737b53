for mode in "light" "heavy"; do
  export GB_FORCE_GENERIC=1; [ $mode = heavy ] && export GB_FORCE_GENERIC_HEAVY=1 || unset GB_FORCE_GENERIC_HEAVY
  for w in headline c2 c4; do
    timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$mode generic $w: %.3e  %.2f ms' % (d['value'], d['ms_per_step']))"
  done
done
unset GB_FORCE_GENERIC GB_FORCE_GENERIC_HEAVY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mockstream.py tests/test_gpu_nbody.py tests/test_gpu_lyapunov.py -q 2>&1 | tail -3
