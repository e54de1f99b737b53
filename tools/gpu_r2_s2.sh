#!/bin/bash
# round 2, GPU session 2: full parity suite, C2 breakdown (per-kernel durations, barrier A/B), C3 pipeline A/B,
# ncu --set full (+ FP64 op counters) of the dominant kernel of every workload.
OUT=gpurun_out/r2s2; mkdir -p $OUT
export GB_PARITY_LOG=$PWD/$OUT/parity_distributions.txt
timeout 1800 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
unset GB_PARITY_LOG
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call"
BENCH_TRACE=1 timeout 600 python bench.py --workload c2 $B > $OUT/c2_default.json 2> $OUT/c2_default.err; echo "c2 default: $(tail -1 $OUT/c2_default.json | cut -c1-140)"; grep trace $OUT/c2_default.err
GB_D8_BLOCKSYNC=0 timeout 600 python bench.py --workload c2 $B > $OUT/c2_nosync.json 2> $OUT/c2_nosync.err; echo "c2 nosync: $(tail -1 $OUT/c2_nosync.json | cut -c1-140)"
timeout 600 python tools/c2_phases.py 303104 1 > $OUT/c2_phases.json 2> $OUT/c2_phases.err; python -c "
import json; d=json.load(open('$OUT/c2_phases.json')); print('phases dense', [{k:round(x,2) for k,x in r.items()} for r in d['rows'][-2:]])"
timeout 600 python tools/c2_phases.py 303104 0 > $OUT/c2_phases_fs.json 2> $OUT/c2_phases_fs.err; python -c "
import json; d=json.load(open('$OUT/c2_phases_fs.json')); print('phases final', [{k:round(x,2) for k,x in r.items()} for r in d['rows'][-2:]])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/launches_c2.log 2>&1
for w in c3 c3d; do
  timeout 600 python bench.py --workload $w $B > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w pipeline: $(tail -1 $OUT/bench_$w.json | cut -c1-140)"; tail -2 $OUT/bench_$w.err
  GB_NO_STREAM_PIPELINE=1 timeout 600 python bench.py --workload $w $B > $OUT/bench_${w}_stepwise.json 2> $OUT/bench_${w}_stepwise.err; echo "$w stepwise: $(tail -1 $OUT/bench_${w}_stepwise.json | cut -c1-140)"
done
FP64M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"
prof() {  # name kernel-regex workload extra-args header
  timeout 900 ncu --set full --metrics $FP64M --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o $OUT/prof_$1 -f python bench.py --workload $3 $4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-single-call > $OUT/ncu_$1.log 2>&1
  python tools/ncu_summary.py $OUT/prof_$1.ncu-rep "$5 (r2s2)" > $OUT/ncu_r2_$1.txt 2> $OUT/ncu_$1.summ.err
  if [ "$6" = src ]; then ncu -i $OUT/prof_$1.ncu-rep --page source --csv > $OUT/prof_$1_source.csv 2>/dev/null; gzip -f $OUT/prof_$1_source.csv; fi
  rm -f $OUT/prof_$1.ncu-rep
  grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|dram__bytes|thread_inst_executed_per|op_dfma" $OUT/ncu_r2_$1.txt | awk '{print "   ", $1, $NF}'
}
prof dop853 k_dop853 c2 "" "ncu --set full --clock-control none, k_dop853_dyn<MW2022, static, dense>, 303,104 orbits x 1000 output times" src
prof leapfrog_mw2022 k_leapfrog headline "" "ncu --set full --clock-control none, k_leapfrog<MW2022, final-state>, 3,031,040 orbits x 1000 steps" src
prof leapfrog_nfw_c1x k_leapfrog c1x "" "ncu --set full --clock-control none, k_leapfrog<NFW, save_all>, 1,048,576 orbits x 256 steps (C1x)"
prof leapfrog_nfw_c1 k_leapfrog c1 "" "ncu --set full --clock-control none, k_leapfrog<NFW, save_all>, 10,000 orbits x 1000 steps (C1)"
prof ruth4_c4 k_ruth4 c4 "" "ncu --set full --clock-control none, k_ruth4<bar+MW2022, rotating, final-state>, 1,212,416 orbits x 1000 steps (C4)"
prof leapfrog_scf_c5 k_leapfrog c5 "" "ncu --set full --clock-control none, k_leapfrog<SCF(10,6), final-state>, 1,250,000 orbits x 1000 steps (C5)"
prof transpose k_transpose c2 "" "ncu --set full --clock-control none, k_transpose_dense<16>, 303,104 orbits x 1000 output times"
ls -la $OUT | head -50
