#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload fusion --steps 3 --warmup 3 > gpurun_out/bench_fusion.json 2> gpurun_out/bench_fusion.err; tail -n 3 gpurun_out/bench_fusion.err; cat gpurun_out/bench_fusion.json
timeout 300 python tools/profile_fusion.py > gpurun_out/prof_fusion.txt 2>&1; tail -n 60 gpurun_out/prof_fusion.txt
