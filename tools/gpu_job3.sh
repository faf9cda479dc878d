#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ctc_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/t_ctc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ctc.log
python tools/perf_kernels.py ctc > gpurun_out/perf_ctc_lin.log 2>&1
tail -15 gpurun_out/t_ctc.log; cat gpurun_out/perf_ctc_lin.log
