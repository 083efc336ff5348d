#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_gemm_tc<\(int\)64>' -s 27 -c 1 -o gpurun_out/conv0_full -f \
    python tools/one_step.py 2 > gpurun_out/ncu_conv0.log 2>&1
tail -1 gpurun_out/ncu_conv0.log
ls -la gpurun_out/conv0_full.ncu-rep
