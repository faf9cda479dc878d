#!/bin/bash
# round-end validation: GPU tests, bench (ours + reference arm), knob matrix, smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu_final.log
tail -n 3 gpurun_out/t_gpu_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -n 2 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}); print(d['e2e']); print(d['hot_path']); print(d['roofline']); print({k:d['beam'][k] for k in ('ms','utt_per_s','e2e_utt_per_s','hbm_frac')}); print(d['ctc']); print(d['cpu_baseline']); print(d['clocks'])
"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 400 gpurun_out/bench_ref_final.json
python tools/knob_matrix.py > gpurun_out/knob_matrix.txt 2>&1; tail -n 1 gpurun_out/knob_matrix.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python tools/exp_ctc_overlap.py 2>&1 | grep -v Warn > gpurun_out/ctc_overlap.txt; tail -n 4 gpurun_out/ctc_overlap.txt
