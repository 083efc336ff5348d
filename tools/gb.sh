python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python tools/gemm_ts.py | cut -c1-200
python tools/gemm_bench.py 2>&1 | cut -c1-150
