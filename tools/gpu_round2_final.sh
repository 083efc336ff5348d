#!/bin/bash
# round-2 ncu evidence (one GPU): launch list of the bench command, DRAM traffic of every hot launch of one step, full sections of
# the three dominant kernels.  Summaries are made from the CSVs afterwards (tools/summarize_ncu.py, tools/traffic_json.py).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-inference > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r2_bench_under_ncu.log | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_lin|k_gemm_tc|k_pattn|k_cross|narrow|k_ln_|k_bn_' --csv --log-file gpurun_out/r2_traffic.csv python tools/one_step.py 2 > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_lin' -s 330 -c 3 -o gpurun_out/r2_lin_full -f python tools/one_step.py 2 > gpurun_out/ncu_full_lin.log 2>&1
tail -1 gpurun_out/ncu_full_lin.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pattn_bwd' -s 9 -c 2 -o gpurun_out/r2_attn_bwd_full -f python tools/one_step.py 2 > gpurun_out/ncu_full_attn.log 2>&1
tail -1 gpurun_out/ncu_full_attn.log
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
