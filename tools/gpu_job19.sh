#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ctc_gpu.py tests/test_trainer_gpu.py -x -q 2>&1 | tail -n 3
python tools/exp_ctc_peaky.py 2>&1 | tail -n 10
python tools/perf_kernels.py ctc 2>&1 | grep -E "T=1000|T=250|T=150"
