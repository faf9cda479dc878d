#!/bin/bash
mkdir -p gpurun_out
python tools/perf_kernels.py ctc > gpurun_out/perf_ctc_lin.log 2>&1
cat gpurun_out/perf_ctc_lin.log
