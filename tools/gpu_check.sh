#!/bin/bash
# One GPU-box visit (1 GPU).  Usage, from the repo root through gpurun: bash tools/gpu_check.sh <step>...
#   tests smoke bench refarm benchwarp benchvote sweep pool wholejob c5 lan malware quick ab ablib stick racecheck
#   experiments ncu nculan ncuvote
# Every step writes under gpurun_out/ and prints one line per result.  No ncu on the large LAN: a full-set replay of a
# kernel that touches 6.7 GB of state per pass does not finish in any budget.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
line() { python -c "import json;d=json.load(open('$1'));print('$2','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['roofline']['kernel'],'checks',(d['checks'] or {}).get('all'))" 2>&1 | tail -1; }
for w in "$@"; do
case "$w" in
tests)
  timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; line gpurun_out/bench.json bench;;
refarm)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; head -c 400 gpurun_out/bench_reference.json; echo;;
benchwarp)
  ITSB_KERNEL=warp timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; line gpurun_out/bench_warp.json "bench warp";;
benchvote)
  ITSB_KERNEL=vote timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vote.json 2> gpurun_out/bench_vote.err; line gpurun_out/bench_vote.json "bench vote";;
sweep)
  # small batches (strong scaling: config C2 split 2, 4, 8 ways), every engine design
  for n in 12500 25000 50000; do for k in warp vote pool; do
    ITSB_KERNEL=$k timeout 600 python bench.py --replicas $n --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/sweep_${k}_${n}.json 2> gpurun_out/sweep_${k}_${n}.err
    line gpurun_out/sweep_${k}_${n}.json "$k $n"
  done; done;;
pool)
  # k_run_pool, every shape, default bench (short)
  for sh in ${POOL_SHAPES:-0 1 2 3 4 5}; do
    ITSB_KERNEL=pool ITSB_POOL_SHAPE=$sh timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/pool_${sh}.json 2> gpurun_out/pool_${sh}.err
    line gpurun_out/pool_${sh}.json "pool shape $sh"
  done;;
wholejob)
  timeout 1200 python tools/whole_job_c2.py > gpurun_out/whole_job_c2.json 2> gpurun_out/whole_job_c2.err; echo "whole job rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/whole_job_c2.json'));print('C2 whole job: %.3e events in %.1f s device, %.1f s wall, failed=%d, oracle-equal replicas %s'%(d['events'],d['device_s'],d['wall_s'],d['first_failed_replica'],sorted(d['replicas_equal_to_the_oracle_at_the_horizon'])))";;
c5)
  for k in ${C5_KERNELS:-warp pool}; do
    ITSB_KERNEL=$k timeout 900 python bench.py --connections 8192 --net-events 1024 --replicas 20000 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c5_${k}.json 2> gpurun_out/c5_${k}.err
    python -c "import json;d=json.load(open('gpurun_out/c5_${k}.json'));print('C5 $k','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['roofline']['kernel'],'d2h/step %.3e'%d['e2e']['d2h_bytes_per_step'],d['config']['records'],'checks',d['checks']['all'])"
  done;;
lan)
  for k in ${LAN_KERNELS:-warp vote pool}; do
    ITSB_KERNEL=$k timeout 600 python bench.py --workload large_lan --replicas 10000 --slice 2 --spinup 6 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/lan_${k}.json 2> gpurun_out/lan_${k}.err
    line gpurun_out/lan_${k}.json "large_lan $k"
  done;;
malware)
  timeout 900 python bench.py --workload malware --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/malware.json 2> gpurun_out/malware.err; line gpurun_out/malware.json malware
  timeout 900 python bench.py --workload backdoor --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/backdoor.json 2> gpurun_out/backdoor.err; line gpurun_out/backdoor.json backdoor;;
quick)
  # after a kernel change: the engine-kernel parity file, then a short default bench
  timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log; tail -3 gpurun_out/pytest_quick.log
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; line gpurun_out/bench_quick.json "bench (8 steps)";;
ab)
  # A/B of two settings of k_run_pool in one visit, alternating (same box, same clocks): ITSB_POOL_BY_PROGRAM 0 / 1
  for rep in 1 2; do for v in 0 1; do
    ITSB_POOL_BY_PROGRAM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ab_${v}_${rep}.json 2> gpurun_out/ab_${v}_${rep}.err
    line gpurun_out/ab_${v}_${rep}.json "by_program=$v run $rep"
  done; done;;
ablib)
  # A/B of two builds in one visit, alternating: the in-tree library vs itsim_b200/libitsim_b200_B.so
  for rep in 1 2 3; do
    timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ablib_A_${rep}.json 2> gpurun_out/ablib_A_${rep}.err
    line gpurun_out/ablib_A_${rep}.json "A run $rep"
    ITSB_LIB=$PWD/itsim_b200/libitsim_b200_B.so timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ablib_B_${rep}.json 2> gpurun_out/ablib_B_${rep}.err
    line gpurun_out/ablib_B_${rep}.json "B run $rep"
  done;;
stick)
  # k_run_pool: warps of an SM staying on one event kind while its ring holds >= N slots (0 = always the fullest ring)
  for n in ${STICKS:-0 8 16 24}; do
    ITSB_POOL_STICK=$n timeout 600 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/stick_${n}.json 2> gpurun_out/stick_${n}.err
    line gpurun_out/stick_${n}.json "pool stick $n"
  done;;
racecheck)
  # the warp-per-replica kernel's "all lanes write the same scalar" model under the race detector: 8 replicas x 300 s
  ITSB_KERNEL=warp timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -c "
import sys; sys.path.insert(0, '.')
from itsim_b200 import netsimple
eng = netsimple.build().run_replicas(8, seed=0)
print('events', eng.run(300.0), eng.kernel_name())
" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck.log;;
experiments)
  # the regrouping variants of round 1 on this round's engine core (the -DITSB_EXPERIMENTS build)
  for k in sorted lanes; do
    ITSB_LIB=$PWD/itsim_b200/libitsim_b200_experiments.so ITSB_KERNEL=$k timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/exp_${k}.json 2> gpurun_out/exp_${k}.err
    line gpurun_out/exp_${k}.json "$k"
  done;;
ncu)
  # launch list of the default bench, then one full capture of its top kernel
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run -s 5 -c 1 -f -o gpurun_out/k_run_default \
     python bench.py --replicas 100000 --steps 2 --warmup 3 --spinup 600 --slice 60 --no-cpu-baseline --no-verify > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la gpurun_out/*.ncu-rep;;
nculan)
  # a SMALL large-LAN capture (4 200 replicas x 0.25 simulated s per launch): where the shipped-itsim path spends its time
  ITSB_KERNEL=pool timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run -s 3 -c 1 -f -o gpurun_out/k_run_pool_hbm \
     python bench.py --workload large_lan --replicas 4200 --slice 0.25 --spinup 4 --steps 2 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/ncu_lan.log 2>&1; echo "ncu lan rc=$?";;
ncuvote)
  ITSB_KERNEL=vote timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run -s 5 -c 1 -f -o gpurun_out/k_run_vote \
     python bench.py --replicas 100000 --steps 2 --warmup 3 --spinup 600 --slice 60 --no-cpu-baseline --no-verify > gpurun_out/ncu_vote.log 2>&1; echo "ncu vote rc=$?";;
*) echo "unknown step $w";;
esac
done
