#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ctc_grad -s 6 -c 1 -o gpurun_out/ctc_grad_r1 -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_grad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_scan_lin -s 6 -c 1 -o gpurun_out/ctc_lin_r1 -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_lin.log 2>&1
tail -3 gpurun_out/ncu_grad.log gpurun_out/ncu_lin.log
