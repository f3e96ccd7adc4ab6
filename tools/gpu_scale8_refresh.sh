#!/bin/bash
# the two headline 8-GPU lines again on the final kernels (config C2 weak scaling, config C3 at full size)
set -u
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus "$@"; }
show() { python -c "import json;d=json.load(open('$1'));print('$1','%.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],d['scaling'],d['n_gpus'],'GPUs',d['roofline']['kernel'],'checks',(d['checks'] or {}).get('all'))" 2>&1 | tail -1; }
NCCL_DEBUG=INFO tr 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_weak_8.json 2> gpurun_out/scale_weak_8.err; show gpurun_out/scale_weak_8.json
grep -m3 "nranks\|NVLS\|comm 0x" gpurun_out/scale_weak_8.err | cut -c1-200
tr 8 --workload malware --scaling strong --total-replicas 1000000 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/c3_8.json 2> gpurun_out/c3_8.err; show gpurun_out/c3_8.json
