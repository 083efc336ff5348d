#!/bin/bash
# ncu evidence for the dominant kernels (one GPU).  Keeps gpurun_out/ small: a metrics-only pass over every launch of two
# steps (DRAM bytes + duration), and `--set full` captures of a handful of launches.
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_gemm_tc|k_patch_attn|k_cross' --csv --log-file gpurun_out/traffic.csv python tools/one_step.py 2 > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log
# full sections: 4 GEMM launches from the middle of the second step (forward block GEMMs), 2 attention backward launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc' -s 360 -c 4 -o gpurun_out/gemm_full -f \
    python tools/one_step.py 2 > gpurun_out/ncu_full_gemm.log 2>&1
tail -1 gpurun_out/ncu_full_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_patch_attn_bwd' -s 9 -c 2 -o gpurun_out/attn_bwd_full -f \
    python tools/one_step.py 2 > gpurun_out/ncu_full_attn.log 2>&1
tail -1 gpurun_out/ncu_full_attn.log
ls -la gpurun_out/*.ncu-rep
