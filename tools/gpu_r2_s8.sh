#!/bin/bash
# round 2, GPU session 8: source-level profile of the SCF leapfrog kernel (C5); sanity of the rebuilt library
OUT=gpurun_out/r2s8; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "c2_dop853 or c5_scf or scf_fortran or known_answers" > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 3 -c 1 -o $OUT/prof_scf -f python bench.py --workload c5 --orbits 303104 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/ncu_scf.log 2>&1
ncu -i $OUT/prof_scf.ncu-rep --page source --csv > $OUT/prof_scf_source.csv 2>/dev/null; gzip -f $OUT/prof_scf_source.csv; rm -f $OUT/prof_scf.ncu-rep; ls -la $OUT
GB_D8_TIMING=1 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/c2.json 2> $OUT/c2.err; echo "c2: $(tail -1 $OUT/c2.json | cut -c1-140)"; tail -1 $OUT/c2.err
