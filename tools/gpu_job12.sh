#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_gpu.py -x -q > gpurun_out/t_lstm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_lstm.log
tail -n 25 gpurun_out/t_lstm.log
timeout 120 python tools/exp_lstm.py 2>&1 | tail -n 10
