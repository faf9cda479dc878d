#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ctc_head_gpu.py -q -x -s > gpurun_out/r2c_head.log 2>&1; echo "head rc=$?" | tee -a gpurun_out/r2c_head.log
grep -v Warning gpurun_out/r2c_head.log | tail -n 15
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2c_all.log
tail -n 6 gpurun_out/r2c_all.log
timeout 300 python bench.py --workload hot --no-comparators > gpurun_out/r2c_hot.json 2> gpurun_out/r2c_hot.err; echo "bench hot rc=$?"
python tools/profile_hot.py > gpurun_out/r2c_hot_timeline.txt 2>&1; echo "profile rc=$?"
