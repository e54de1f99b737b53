#!/bin/bash
# One GPU-box session: parity tests, the bench lines, launch list and ncu captures -> gpurun_out/<tag>/
# usage: tools/gpu_session.sh <tag> [steps...]   steps: tests bench ncu_lf ncu_d8 launches
TAG=${1:-s}; shift
STEPS=${@:-tests bench ncu_lf ncu_d8 launches}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
# the .ncu-rep files (with source) exceed gpurun's 64 MiB return limit: summarise on the box, keep the text
summ() { python tools/ncu_summary.py $OUT/$1.ncu-rep "$2" > $OUT/$1.txt 2> $OUT/$1.summ.err; ls -la $OUT/$1.ncu-rep | awk '{print "rep bytes", $5}'; [ -z "$KEEP_REP" ] && rm -f $OUT/$1.ncu-rep; grep -E "gpu__time_duration|pipe_fp64_cycles_active|registers_per_thread|warps_active|dram__bytes" $OUT/$1.txt | head -8; }
for s in $STEPS; do
case $s in
tests)   timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log ;;
bench)   for w in headline c1 c2 c4 c5; do
            extra="--no-cpu-baseline"; [ $w = headline ] && extra=""
            timeout 900 python bench.py --workload $w --steps 3 --warmup 3 $extra > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -1 $OUT/bench_$w.json | cut -c1-400
         done ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1 ;;
ncu_lf)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 3 -c 1 -o $OUT/prof_leapfrog -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_lf.log 2>&1; summ prof_leapfrog "ncu --set full --clock-control none, k_leapfrog<MW2022, final-state>, bench headline size ($TAG)" ;;
ncu_d8)  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_dop853 -s 3 -c 1 -o $OUT/prof_dop853 -f python bench.py --workload c2 --orbits ${D8_ORBITS:-75776} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_d8.log 2>&1; summ prof_dop853 "ncu --set full --clock-control none, k_dop853_dyn<MW2022, static, dense>, 75,776 orbits x 1000 output times ($TAG)" ;;
ncu_r4)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ruth4 -s 3 -c 1 -o $OUT/prof_ruth4 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_r4.log 2>&1; summ prof_ruth4 "ncu --set full --clock-control none, k_ruth4<bar+MW2022, rotating, final-state>, C4 bench size ($TAG)" ;;
ncu_scf) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog -s 3 -c 1 -o $OUT/prof_scf -f python bench.py --workload c5 --orbits 303104 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_scf.log 2>&1; summ prof_scf "ncu --set full --clock-control none, k_leapfrog<SCF(10,6), final-state>, 303,104 orbits x 1000 steps ($TAG)" ;;
c2ab)    for v in sort nosort; do
            [ $v = nosort ] && export GB_D8_NOSORT=1 || unset GB_D8_NOSORT
            timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_$v.json 2> $OUT/bench_c2_$v.err; tail -1 $OUT/bench_c2_$v.json | cut -c1-300
         done; unset GB_D8_NOSORT ;;
c2ni)    for v in inline noinline; do
            [ $v = noinline ] && export GALA_B200_LIB=$PWD/gala_b200/libgala_b200_noinline.so || unset GALA_B200_LIB
            timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_$v.json 2> $OUT/bench_c2_$v.err; tail -1 $OUT/bench_c2_$v.json | cut -c1-300
         done; unset GALA_B200_LIB ;;
c2bs)    for v in "64 0" "128 1" "256 1" "128 0"; do
            set -- $v; export GB_D8_BLOCK=$1 GB_D8_BLOCKSYNC=$2
            timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_c2_b$1_s$2.json 2> $OUT/bench_c2_b$1_s$2.err; echo "block $1 sync $2: $(tail -1 $OUT/bench_c2_b$1_s$2.json | cut -c1-120)"
         done; unset GB_D8_BLOCK GB_D8_BLOCKSYNC ;;
launches_c2) ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_c2.log 2>&1; grep -c k_ $OUT/launches_c2.csv ;;
c3)      for w in c3 c3d; do timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -1 $OUT/bench_$w.json | cut -c1-200; tail -2 $OUT/bench_$w.err; done ;;
b1)      w=$B1W; timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -1 $OUT/bench_$w.json | cut -c1-220; tail -2 $OUT/bench_$w.err ;;
d8stats) timeout 600 python tools/d8_stats.py > $OUT/d8_stats.json 2> $OUT/d8_stats.err; head -50 $OUT/d8_stats.json ;;
d8ab)    for v in default loop loop2 noinline; do
            [ $v = default ] && unset GALA_B200_LIB || export GALA_B200_LIB=$PWD/gala_b200/libgala_b200_$v.so
            timeout 600 python tools/c2_phases.py 303104 1 > $OUT/phases_$v.json 2> $OUT/phases_$v.err; python -c "
import json,sys; d=json.load(open('$OUT/phases_$v.json')); r=d['rows'][-1]; print('$v dense', {k:round(x,2) for k,x in r.items()})"
            timeout 600 python tools/c2_phases.py 303104 0 > $OUT/phases_fs_$v.json 2> $OUT/phases_fs_$v.err; python -c "
import json,sys; d=json.load(open('$OUT/phases_fs_$v.json')); r=d['rows'][-1]; print('$v final', {k:round(x,2) for k,x in r.items()})"
            [ $v != default ] && { timeout 900 python -m pytest tests -m gpu -x -q -k "dop853 or mockstream" > $OUT/pytest_$v.log 2>&1; tail -1 $OUT/pytest_$v.log; }
         done; unset GALA_B200_LIB ;;
*) echo "unknown step $s" ;;
esac
done
