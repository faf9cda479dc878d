#!/bin/bash
# round 2 closing evidence pass: full GPU suite, the driver's bench command, smoke, beam A/B, ncu of the beam kernels
# (the knob matrix ran in an earlier pass of the same script generation: profiles/r02_knob_matrix.txt)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2g_all.log
tail -n 3 gpurun_out/r2g_all.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}); print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic']); print({k:d['beam'][k] for k in ('ms','hbm_frac','route','shard_sweep')}); print(d['clocks'])
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 100 python tools/exp_beam_fused.py gpurun_out/r2g_beam_ab.txt > /dev/null 2>&1; cat gpurun_out/r2g_beam_ab.txt | grep -E "N= 4096|N= 2048|N= 1024|N=  512|N=   16" | grep -v auto
timeout 150 ncu --set full --clock-control none --import-source on -k regex:beam_ --launch-skip 2 -c 3 -o gpurun_out/r2g_beam -f python tools/run_beam_once.py > gpurun_out/r2g_ncu_beam.log 2>&1; echo "ncu beam rc=$?"
