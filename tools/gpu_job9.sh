#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 40 -c 3 -o gpurun_out/gemm_r1 -f \
    python tools/profile_fusion.py > gpurun_out/ncu_gemm.log 2>&1
tail -n 3 gpurun_out/ncu_gemm.log
