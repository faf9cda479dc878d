#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_beam_gpu.py -x -q > gpurun_out/t_beam.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_beam.log
tail -n 3 gpurun_out/t_beam.log
timeout 300 python bench.py --workload beam --steps 3 --warmup 3 > gpurun_out/bench_beam.json 2> gpurun_out/bench_beam.err; tail -n 2 gpurun_out/bench_beam.err; python -c "
import json; d=json.load(open('gpurun_out/bench_beam.json')); print({k:d['beam'][k] for k in ('ms','utt_per_s','e2e_ms','e2e_utt_per_s','hbm_frac')})"
