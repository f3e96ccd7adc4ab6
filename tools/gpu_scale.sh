#!/bin/bash
# Multi-GPU lines (run on a box with >= $1 GPUs): weak and strong scaling of config C2, config C3 at full size, config C5.
# usage: bash tools/gpu_scale.sh <max_gpus> [quick]
set -u
MAX=${1:-2}; QUICK=${2:-}
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus "$@"; }
show() { python -c "import json,sys;d=json.load(open('$1'));print('$1','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['scaling'],d['n_gpus'],'GPUs',d['roofline']['kernel'],d['config'].get('processes'),'allreduce_ms',d['config']['telemetry_allreduce_ms'],'checks',(d['checks'] or {}).get('all'))" 2>&1 | tail -1; }
for N in ${SCALE_NS:-2 4 8}; do
  [ "$N" -gt "$MAX" ] && continue
  NCCL_DEBUG=WARN tr $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_weak_${N}.json 2> gpurun_out/scale_weak_${N}.err; show gpurun_out/scale_weak_${N}.json
  tr $N --steps 8 --warmup 3 --no-cpu-baseline --scaling strong --total-replicas 100000 > gpurun_out/scale_strong_${N}.json 2> gpurun_out/scale_strong_${N}.err; show gpurun_out/scale_strong_${N}.json
done
# one process, one host thread per GPU (no launcher)
python bench.py --gpus $MAX --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_threads_${MAX}.json 2> gpurun_out/scale_threads_${MAX}.err; show gpurun_out/scale_threads_${MAX}.json
[ -n "$QUICK" ] && exit 0
# config C3 at full size: 1 000 000 malware replicas over the GPUs, histograms over all replicas reduced with NCCL
tr $MAX --workload malware --scaling strong --total-replicas 1000000 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/c3_${MAX}.json 2> gpurun_out/c3_${MAX}.err; show gpurun_out/c3_${MAX}.json
# config C5: full logging, every record drained to the host every step
python bench.py --connections 8192 --net-events 1024 --replicas 20000 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c5_1.json 2> gpurun_out/c5_1.err; show gpurun_out/c5_1.json
for N in 2 4 8; do
  [ "$N" -gt "$MAX" ] && continue
  tr $N --connections 8192 --net-events 1024 --replicas 20000 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/c5_${N}.json 2> gpurun_out/c5_${N}.err; show gpurun_out/c5_${N}.json
done
