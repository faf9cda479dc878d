#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_infonce_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/t_nce.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_nce.log
python tools/profile_hot.py hot > gpurun_out/prof_hot2.log 2>&1
python tools/exp_visual.py > gpurun_out/exp_visual.log 2>&1
tail -3 gpurun_out/t_nce.log; head -8 gpurun_out/exp_visual.log
