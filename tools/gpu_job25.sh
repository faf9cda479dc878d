#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_ctc_once.py > gpurun_out/run_ctc_once.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ctc_scan_ws -s 2 -c 1 -o gpurun_out/ctc_ws_r1d -f python tools/run_ctc_once.py > gpurun_out/ncu_ws.log 2>&1
tail -n 3 gpurun_out/ncu_ws.log
