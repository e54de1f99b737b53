#!/bin/bash
# round 2, GPU session 10: DRAM traffic of the dominant kernel of the mock-stream workloads (largest launch of each kernel)
OUT=gpurun_out/r2s10; mkdir -p $OUT
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-single-call"
for v in "c3d k_mock_dop853" "c3sg k_nbody_leapfrog" "c3sgd k_nbody_dop853"; do set -- $v
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$2 -c 12 --csv --log-file $OUT/traffic_$1.csv python bench.py --workload $1 $B > $OUT/ncu_$1.log 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/traffic_$1.csv")))
hdr = [r for r in rows if "Metric Name" in r][0]
iid, im, iv, iu = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.defaultdict(dict)
for r in rows:
    if len(r) == len(hdr) and r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        v = float(r[iv].replace(",", "")); u = r[iu]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1)
        per[r[iid]][r[im]] = v
best = max(per.values(), key=lambda d: d.get("gpu__time_duration.sum", 0))
print("$1", "$2", "launches", len(per), "largest:", best)
PY
done
for w in c3 c3d c3sg c3sgd; do :; done
