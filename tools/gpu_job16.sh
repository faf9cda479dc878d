#!/bin/bash
mkdir -p gpurun_out
python tools/exp_visual2.py 2>&1 | cut -c1-150 | grep -v "^-" | head -20
python -m pytest tests/test_trainer_gpu.py -x -q 2>&1 | tail -n 3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -n 3 gpurun_out/bench4.err
python -c "
import json; d=json.load(open('gpurun_out/bench4.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['hot_path']); print(d['roofline'])
"
