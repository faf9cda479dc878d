#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/t_lstm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_lstm.log
tail -n 3 gpurun_out/t_lstm.log
timeout 600 python tools/profile_hot.py hot > gpurun_out/prof_hot5.log 2>&1; grep -n "wall ms\|GPU busy\|lstm_cast" gpurun_out/prof_hot5.log | head -5
