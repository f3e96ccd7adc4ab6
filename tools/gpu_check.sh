#!/bin/bash
# One GPU-box visit: parity suite, the default bench, the ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, through gpurun): bash tools/gpu_check.sh [tests|bench|ncu]...   (default: all three)
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | head -c 1500; echo;;
benchwarp)
  ITSB_KERNEL=warp timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; echo "bench warp rc=$?"; head -c 600 gpurun_out/bench_warp.json; echo;;
sweep)
  # small batches (strong scaling: config C2 split 2, 4, 8 ways), both engine designs
  for n in 12500 25000 50000; do for k in warp vote; do
    ITSB_KERNEL=$k timeout 600 python bench.py --replicas $n --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/sweep_${k}_${n}.json 2> gpurun_out/sweep_${k}_${n}.err
    python -c "import json;d=json.load(open('gpurun_out/sweep_${k}_${n}.json'));print('$k',$n,'%.3e'%d['value'],d['roofline']['kernel'])"
  done; done;;
pool)
  # k_run_pool, every shape, default bench (short)
  for sh in ${POOL_SHAPES:-0 1 2 3 4 5}; do
    ITSB_KERNEL=pool ITSB_POOL_SHAPE=$sh timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/pool_${sh}.json 2> gpurun_out/pool_${sh}.err
    python -c "import json;d=json.load(open('gpurun_out/pool_${sh}.json'));print('pool shape $sh','%.3e'%d['value'],d['roofline']['kernel'])"
  done;;
ncupool)
  ITSB_KERNEL=pool timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_run -s 5 -c 1 -f -o gpurun_out/k_run_pool \
     python bench.py --replicas 100000 --steps 2 --warmup 3 --spinup 600 --slice 60 --no-cpu-baseline --no-verify > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?";;
ncuwholejob)
  timeout 1200 python tools/whole_job_c2.py > gpurun_out/whole_job_c2.json 2> gpurun_out/whole_job_c2.err; echo "whole job rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/whole_job_c2.json'));print('C2 whole job: %.3e events in %.1f s device, %.1f s wall, failed=%d, oracle-equal replicas %s'%(d['events'],d['device_s'],d['wall_s'],d['first_failed_replica'],sorted(d['replicas_equal_to_the_oracle_at_the_horizon'])))";;
c5)
  for k in warp pool; do
    ITSB_KERNEL=$k timeout 900 python bench.py --connections 8192 --net-events 1024 --replicas 20000 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c5_${k}.json 2> gpurun_out/c5_${k}.err
    python -c "import json;d=json.load(open('gpurun_out/c5_${k}.json'));print('C5 $k','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['roofline']['kernel'],'d2h/step %.3e'%d['e2e']['d2h_bytes_per_step'],d['config']['records'],'checks',d['checks']['all'])"
  done;;
lan)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_run -s 2 -c 1 -f -o gpurun_out/k_run_vote_hbm \
     python bench.py --workload large_lan --replicas 10000 --slice 2 --spinup 6 --steps 2 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/ncu_lan.log 2>&1; echo "ncu lan rc=$?";;
poolsweep)
  for n in 12500 25000 50000 200000; do
    ITSB_KERNEL=pool timeout 600 python bench.py --replicas $n --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/sweep_pool_${n}.json 2> gpurun_out/sweep_pool_${n}.err
    python -c "import json;d=json.load(open('gpurun_out/sweep_pool_${n}.json'));print('pool',$n,'%.3e'%d['value'],d['roofline']['kernel'])"
  done;;
wholejob)
  timeout 1200 python tools/whole_job_c2.py > gpurun_out/whole_job_c2.json 2> gpurun_out/whole_job_c2.err; echo "whole job rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/whole_job_c2.json'));print('C2 whole job: %.3e events in %.1f s device, %.1f s wall, failed=%d, oracle-equal replicas %s'%(d['events'],d['device_s'],d['wall_s'],d['first_failed_replica'],sorted(d['replicas_equal_to_the_oracle_at_the_horizon'])))";;
c5)
  for k in warp pool; do
    ITSB_KERNEL=$k timeout 900 python bench.py --connections 8192 --net-events 1024 --replicas 20000 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c5_${k}.json 2> gpurun_out/c5_${k}.err
    python -c "import json;d=json.load(open('gpurun_out/c5_${k}.json'));print('C5 $k','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['roofline']['kernel'],'d2h/step %.3e'%d['e2e']['d2h_bytes_per_step'],d['config']['records'],'checks',d['checks']['all'])"
  done;;
lan)
  for k in ${LAN_KERNELS:-warp vote pool}; do
    ITSB_KERNEL=$k timeout 900 python bench.py --workload large_lan --replicas 10000 --slice 2 --spinup 6 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/lan_${k}.json 2> gpurun_out/lan_${k}.err
    python -c "import json;d=json.load(open('gpurun_out/lan_${k}.json'));print('large_lan $k','%.3e'%d['value'],d['roofline']['kernel'])"
  done;;
experiments)
  # the regrouping variants of round 1 on this round's engine core (the -DITSB_EXPERIMENTS build)
  for k in pool sorted lanes; do
    ITSB_LIB=$PWD/itsim_b200/libitsim_b200_experiments.so ITSB_KERNEL=$k timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/exp_${k}.json 2> gpurun_out/exp_${k}.err
    python -c "import json;d=json.load(open('gpurun_out/exp_${k}.json'));print('$k','%.3e'%d['value'],d['roofline']['kernel'])"
  done;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_run -s 5 -c 1 -f -o gpurun_out/k_run_vote \
     python bench.py --replicas 100000 --steps 2 --warmup 3 --spinup 600 --slice 60 --no-cpu-baseline --no-verify > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la gpurun_out/*.ncu-rep;;
esac
done
