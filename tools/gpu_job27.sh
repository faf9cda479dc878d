#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_lstm_gpu.py -x -q > gpurun_out/t_tr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_tr.log
tail -n 6 gpurun_out/t_tr.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -n 3 gpurun_out/bench_train.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','hot_path','gpu_launches')})
PY
