#!/bin/bash
# dev: [full gpu tests] + one bench run + per-call native profile.  usage: tools/gpu_step.sh [notest]
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
fi
LOTUS_BENCH_DUMP=gpurun_out/step_calls.json timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
python -c "
import json
d=json.load(open('gpurun_out/bench_s.json')); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), round(d['roofline']['avg_launch_us'],1), 'attn', round(d['attention_block']['frac'],3), 'prof_ms', round(d['kernel_ms_per_step_profiled'],2))
for k in d['kernels'][:24]: print('  %-40s %4d %7.3f %s' % (k['call'], k['calls'], k['ms_per_step'], k['frac_of_roofline']))
" || tail -5 gpurun_out/bench_s.err
