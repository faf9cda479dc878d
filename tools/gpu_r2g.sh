#!/bin/bash
# round 2 closing evidence pass: full GPU suite, the driver's bench command, smoke, knob matrix
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2g_all.log
tail -n 4 gpurun_out/r2g_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}); print(d['e2e']); print(d['roofline']); print(d['beam']); print(d['clocks'])
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python tools/knob_matrix.py > gpurun_out/r2g_knob_matrix.txt 2>&1; grep "###" gpurun_out/r2g_knob_matrix.txt
