#!/bin/bash
# training bench at N GPUs of one node (the driver's launch line); usage: tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 24 --warmup 5 --no-inference --no-cpu-baseline 2> gpurun_out/r2_bench_${N}gpu.err | grep '^{' > gpurun_out/r2_bench_${N}gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json')); print('gpus', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/r2_bench_${N}gpu.err
