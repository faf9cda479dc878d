#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload ctc --steps 3 --warmup 3 > gpurun_out/bench_ctc2.json 2> gpurun_out/bench_ctc2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ctc2.csv \
    python bench.py --workload ctc --steps 3 --warmup 3 > gpurun_out/ncu_launch2.log 2>&1
python tools/perf_kernels.py ctc1000 > gpurun_out/plain_ctc1000b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ctc_scan_ws|ctc_grad_lin" -s 12 -c 2 -o gpurun_out/ctc_r1c -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_full2.log 2>&1
python tools/perf_kernels.py beam1 > gpurun_out/plain_beam1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"beam_" -s 4 -c 2 -o gpurun_out/beam_r1 -f \
    python tools/perf_kernels.py beam1 > gpurun_out/ncu_beam.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 9000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_train.log 2>&1
ls -la gpurun_out | tail -n 12
