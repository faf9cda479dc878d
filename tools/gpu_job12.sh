#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_beam_gpu.py -x -q > gpurun_out/t_beam.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_beam.log
tail -n 5 gpurun_out/t_beam.log
timeout 300 python bench.py --workload beam --steps 3 --warmup 3 > gpurun_out/bench_beam.json 2> gpurun_out/bench_beam.err; tail -n 3 gpurun_out/bench_beam.err; cat gpurun_out/bench_beam.json
