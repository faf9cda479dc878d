#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ctc_scan_ws -s 7 -c 1 -o gpurun_out/ctc_ws_r1 -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_ws.log 2>&1
tail -n 3 gpurun_out/ncu_ws.log
