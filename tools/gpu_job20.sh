#!/bin/bash
python -m pytest tests/test_trainer_gpu.py -x -q 2>&1 | tail -n 2
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -n 2 gpurun_out/bench5.err
python -c "
import json; d=json.load(open('gpurun_out/bench5.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print(d['e2e'])
"
