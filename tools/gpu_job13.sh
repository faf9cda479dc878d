#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
tail -n 4 gpurun_out/t_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -n 3 gpurun_out/bench3.err
python -c "
import json; d=json.load(open('gpurun_out/bench3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['hot_path']); print(d['roofline']); print(d['fusion']); print({k:d['beam'][k] for k in ('ms','utt_per_s','e2e_utt_per_s','hbm_frac')}); print(d['ctc'])
"
python tools/profile_hot.py hot > gpurun_out/prof_hot3.log 2>&1
