#!/bin/bash
# round-1 GPU job: tests, full-step profile, ncu launch list + full capture of the CTC kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
python tools/profile_hot.py full > gpurun_out/prof_full.log 2>&1
python bench.py --workload ctc --steps 3 --warmup 3 > gpurun_out/bench_ctc.json 2> gpurun_out/bench_ctc.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ctc.csv \
    python bench.py --workload ctc --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
python tools/perf_kernels.py ctc1000 > gpurun_out/plain_ctc1000.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_ -s 8 -c 4 -o gpurun_out/ctc_r1b -f \
    python tools/perf_kernels.py ctc1000 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_gpu.log
