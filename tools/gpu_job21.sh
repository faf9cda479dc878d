#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_lstm_gpu.py -x -q > gpurun_out/t_lstm2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_lstm2.log
tail -n 12 gpurun_out/t_lstm2.log
timeout 120 python tools/exp_lstm.py 2>&1 | tail -n 8
