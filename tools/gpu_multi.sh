#!/bin/bash
# N-GPU check of the data-parallel step (one process per GPU, NCCL gradient all-reduce) + the reference (CPU) arm
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-inference > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${N}gpu.json"))
    print("bench ${N} GPUs", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["n_gpus"], d["config"]["parallelism"])
except Exception as e:
    print("failed", e)
PY
