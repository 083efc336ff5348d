#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("bench", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["roofline"]["frac"], d["attention_block"]["frac"], d["inference"]["p50_ms_per_keystep"], d["cpu_baseline"]["value"])
except Exception as e:
    print("bench failed", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py 3 > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log
