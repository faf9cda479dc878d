#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_beam_gpu.py -x -q > gpurun_out/t_beam.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_beam.log
tail -n 3 gpurun_out/t_beam.log
timeout 600 python -m pytest tests/test_ctc_gpu.py -x -q > gpurun_out/t_ctc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ctc.log
tail -n 4 gpurun_out/t_ctc.log
timeout 300 python tools/exp_pf.py > gpurun_out/exp_pf.txt 2>&1; tail -n 3 gpurun_out/exp_pf.txt
