#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_gemm_gpu.py tests/test_trainer_gpu.py tests/test_lstm_gpu.py -x -q > gpurun_out/t_fus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_fus.log
tail -n 8 gpurun_out/t_fus.log
timeout 300 python bench.py --workload fusion --steps 3 --warmup 3 > gpurun_out/bench_fusion.json 2> gpurun_out/bench_fusion.err; tail -n 3 gpurun_out/bench_fusion.err; cat gpurun_out/bench_fusion.json
