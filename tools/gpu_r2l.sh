#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ctc_gpu.py tests/test_bench_sizes_gpu.py tests/test_lstm_gpu.py -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2l_tests.log
tail -n 4 gpurun_out/r2l_tests.log
timeout 300 python bench.py --workload ctc --no-comparators > gpurun_out/r2l_ctc.json 2> gpurun_out/r2l_ctc.err; echo "bench ctc rc=$?"
python - <<'PY'
import json
c=json.load(open('gpurun_out/r2l_ctc.json'))['ctc']
print({T:{k:round(c[T][k],4) for k in ('product_ms','product_eager_ms','abi_fwd_bwd_ms','scan_ms','grad_ms','gbs')} for T in c})
PY
timeout 600 python bench.py --steps 10 --warmup 4 --no-comparators --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step')}); print(d['e2e']); print(d['hot_path']['ms_per_step'])
PY
