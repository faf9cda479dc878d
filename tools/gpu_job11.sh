#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_fusion_gpu.py -x -q > gpurun_out/t_tr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_tr.log
tail -n 4 gpurun_out/t_tr.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -n 3 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
