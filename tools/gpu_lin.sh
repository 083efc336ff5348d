#!/bin/bash
# dev: tensor-core kernel tests + linear micro-benchmark
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_tc.py -m gpu -q -x ) > gpurun_out/pytest_tc.log 2>&1
tail -5 gpurun_out/pytest_tc.log
( timeout 100 python tools/gemm_bench.py ) > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log | tail -20
