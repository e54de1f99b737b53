#!/bin/bash
# round 2, GPU session 6: TimeInterpolated tests incl. the mirrored rotating-bar test; ncu traffic of the mock-stream kernels
OUT=gpurun_out/r2s6; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_timeinterp.py tests/test_gpu_parity.py -m gpu -q -s -k "timeinterp or rotating_bar or evaluation_parity or integration_parity or docstring or unsupported or step_statistics" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log; grep -E "rotating bar|timeinterp " $OUT/pytest.log | head -12
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
for v in "c3 k_mock_leapfrog 8" "c3d k_mock_dop853 1" "c3sg k_nbody_leapfrog 1" "c3sgd k_nbody_dop853 1"; do set -- $v
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$2 -s $((3 * $3)) -c $3 --csv --log-file $OUT/traffic_$1.csv python bench.py --workload $1 $B > $OUT/ncu_$1.log 2>&1
  python - <<PY
import csv
rows = list(csv.reader(open("$OUT/traffic_$1.csv")))
hdr = [r for r in rows if "Metric Name" in r][0]
im, iv, iu = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows:
    if len(r) == len(hdr) and r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        v = float(r[iv].replace(",", "")); u = r[iu]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(u, 1)
        tot[r[im]] = tot.get(r[im], 0) + v
print("$1", "$2", tot)
PY
done
