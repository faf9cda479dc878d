#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_ctc_once.py > gpurun_out/run_ctc_once.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:ctc_scan_ws|ctc_grad_lin' -s 2 -c 2 -o gpurun_out/ctc_r1e -f python tools/run_ctc_once.py > gpurun_out/ncu_ctc_e.log 2>&1
tail -n 2 gpurun_out/ncu_ctc_e.log
timeout 600 python bench.py --workload ctc --steps 2 --warmup 3 > gpurun_out/bench_ctc3.json 2> gpurun_out/bench_ctc3.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ctc3.csv python bench.py --workload ctc --steps 2 --warmup 3 > gpurun_out/ncu_launch3.log 2>&1
cat gpurun_out/bench_ctc3.json | cut -c1-600
