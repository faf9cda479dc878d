#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_beam_gpu.py -x -q > gpurun_out/t_beam.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/t_beam.log
tail -n 5 gpurun_out/t_beam.log
timeout 300 python bench.py --workload beam --steps 3 --warmup 3 > gpurun_out/bench_beam.json 2> gpurun_out/bench_beam.err; brc=$?; tail -n 3 gpurun_out/bench_beam.err; cat gpurun_out/bench_beam.json
if [ $rc -eq 0 ] && [ $brc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:beam_ -c 2 -o gpurun_out/beam_r1c -f python bench.py --workload beam --steps 1 --warmup 1 > gpurun_out/ncu_beam.log 2>&1
  tail -n 3 gpurun_out/ncu_beam.log
fi
