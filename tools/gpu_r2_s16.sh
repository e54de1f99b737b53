#!/bin/bash
# round 2, GPU session 16: one body + one particle as a pair of lanes (k_nbody_dop853_pair) vs the one-lane n = 12 form
OUT=gpurun_out/r2s16; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_nbody.py tests/test_gpu_mockstream.py tests/test_gpu_parity.py -m gpu -q -s > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log; grep -E "self-gravity|nbody dop853" $OUT/pytest.log | head
GB_NBODY_NO_PAIR=1 timeout 600 python -m pytest tests/test_gpu_nbody.py -m gpu -q > $OUT/pytest_nopair.log 2>&1; tail -1 $OUT/pytest_nopair.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
timeout 600 python bench.py --workload c3sgd $B > $OUT/c3sgd_pair.json 2> $OUT/c3sgd_pair.err; echo "c3sgd pair: $(tail -1 $OUT/c3sgd_pair.json | cut -c1-150)"
GB_NBODY_NO_PAIR=1 timeout 600 python bench.py --workload c3sgd $B > $OUT/c3sgd_onelane.json 2> $OUT/c3sgd_onelane.err; echo "c3sgd one lane: $(tail -1 $OUT/c3sgd_onelane.json | cut -c1-150)"
timeout 600 python bench.py --workload c3sg $B > $OUT/c3sg.json 2> $OUT/c3sg.err; echo "c3sg: $(tail -1 $OUT/c3sg.json | cut -c1-150)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_nbody_dop853 -c 12 --csv --log-file $OUT/traffic_c3sgd.csv python bench.py --workload c3sgd --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-single-call > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/traffic_c3sgd.csv")))
hdr = [r for r in rows if "Metric Name" in r][0]
iid, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.defaultdict(dict)
for r in rows:
    if len(r) == len(hdr) and r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        v = float(r[iv].replace(",", "")); u = r[iu]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1)
        per[r[iid]][r[im]] = v; per[r[iid]]["k"] = r[ik][:40]
best = max(per.values(), key=lambda d: d.get("gpu__time_duration.sum", 0))
print("c3sgd largest launch:", best)
PY
