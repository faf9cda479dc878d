#!/bin/bash
# round 2, first GPU pass: all GPU tests, the bench line (ours + reference arm), one ncu --set full capture of the CTC pair
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -n 5 gpurun_out/r2a_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
python tools/run_ctc_once.py > gpurun_out/r2a_ctc_once.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_ -o gpurun_out/r2a_ctc python tools/run_ctc_once.py > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -n 8
