#!/bin/bash
mkdir -p gpurun_out
python tools/perf_kernels.py ctc1000 > gpurun_out/plain_ctc1000c.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ctc_grad_lin" -s 6 -c 1 -o gpurun_out/ctc_gradlin_r1 -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_full3.log 2>&1
tail -n 2 gpurun_out/ncu_full3.log
