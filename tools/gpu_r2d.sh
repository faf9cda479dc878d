#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graphs_gpu.py tests/test_ctc_head_gpu.py tests/test_trainer_gpu.py tests/test_bench_sizes_gpu.py -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2d_tests.log
tail -n 25 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 10 --warmup 4 --no-comparators --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['hot_path'])
PY
