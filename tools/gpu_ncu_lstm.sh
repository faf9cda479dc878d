#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/run_lstm_once.py > gpurun_out/run_lstm_once.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_kernel\|lstm_bwd_kernel -s 4 -c 2 -o gpurun_out/lstm_r1 -f python tools/run_lstm_once.py > gpurun_out/ncu_lstm.log 2>&1
tail -n 3 gpurun_out/ncu_lstm.log
