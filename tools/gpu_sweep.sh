#!/bin/bash
# inference sweep (BASELINE configs[4]) at N replicas; usage: tools/gpu_sweep.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 900 python tools/infer_sweep.py --batches 1,4,16,64,256 --iters 10 --warmup 3 > gpurun_out/r2_infer_sweep_${N}gpu.json 2> gpurun_out/r2_infer_sweep_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/infer_sweep.py --batches 1,4,16,64,256 --iters 10 --warmup 3 > gpurun_out/r2_infer_sweep_${N}gpu.json 2> gpurun_out/r2_infer_sweep_${N}gpu.err
fi
tail -3 gpurun_out/r2_infer_sweep_${N}gpu.err | cut -c1-400
python -c "
import json
d=json.load(open('gpurun_out/r2_infer_sweep_${N}gpu.json'))
for c in d['cells']: print(c.get('points'), c.get('batch'), round(c.get('p50_ms',0),2), round(c.get('keysteps_per_s',0)), c.get('frac_of_roofline') and round(c['frac_of_roofline'],3), c.get('skipped',''))
"
