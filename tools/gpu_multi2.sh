#!/bin/bash
# dev: 2-GPU bench, NCCL channel-count sweep with the overlapped all-reduce
mkdir -p gpurun_out
for ch in 2 4 8 16; do
  NCCL_MAX_NCHANNELS=$ch timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-inference > gpurun_out/bench_2g_ch$ch.json 2> gpurun_out/bench_2g_ch$ch.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_2g_ch$ch.json')); print('channels $ch', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/bench_2g_ch$ch.err
done
