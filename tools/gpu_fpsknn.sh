#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fps or knn or ball or tokenis" 2>&1 | tail -2
timeout 300 python tools/fps_knn_bench.py > gpurun_out/r2_fps_knn.json 2> gpurun_out/r2_fps_knn.err; tail -c 1500 gpurun_out/r2_fps_knn.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_fps|k_knn|k_ball|k_cloud_sort|k_hash|k_codes|k_pair|k_patch_plan|k_pool" -c 60 --csv --log-file gpurun_out/r2_ncu_fps_knn.csv python tools/fps_knn_bench.py > /dev/null 2>&1
wc -l gpurun_out/r2_ncu_fps_knn.csv
