#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_beam_gpu.py -x -q > gpurun_out/t_beam.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_beam.log
python tools/perf_kernels.py beam > gpurun_out/perf_beam.log 2>&1
tail -n 5 gpurun_out/t_beam.log; cat gpurun_out/perf_beam.log
