#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x ) > gpurun_out/pytest_tc.log 2>&1
tail -2 gpurun_out/pytest_tc.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err
python -c "
import json
d=json.load(open('gpurun_out/bench_q$i.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['avg_launch_us'])" || tail -3 gpurun_out/bench_q$i.err
done
