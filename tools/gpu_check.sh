#!/bin/bash
# GPU-box check: parity tests, bench line (A/B of the weight-gradient stream); `full` adds the ncu launch list.
# (keep gpurun_out/ under 64 MiB: full ncu captures are done separately, a few launches at a time)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
LOTUS_NO_SIDE_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_noside.json 2> gpurun_out/bench_noside.err
python - <<'PY'
import json
for f in ("bench", "bench_noside"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("inference"))
    except Exception as e:
        print(f, "failed", e)
PY
if [ "$1" == "full" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/one_step.py 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
fi
