#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_beam_gpu.py tests/test_ctc_gpu.py -x -q > gpurun_out/t_bc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_bc.log
tail -n 4 gpurun_out/t_bc.log
timeout 600 python tools/exp_pf.py > gpurun_out/exp_pf.txt 2>&1; tail -n 6 gpurun_out/exp_pf.txt
