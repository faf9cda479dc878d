#!/bin/bash
# round 2, fused attention + grouped GEMM bring-up: dedicated attention test first (under a timeout: a hang must not strike)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x -s > gpurun_out/r2b_att.log 2>&1; echo "att rc=$?" | tee -a gpurun_out/r2b_att.log
tail -n 15 gpurun_out/r2b_att.log
timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_gemm_gpu.py tests/test_bench_sizes_gpu.py tests/test_trainer_gpu.py -q -x -s > gpurun_out/r2b_fus.log 2>&1; echo "fusion rc=$?" | tee -a gpurun_out/r2b_fus.log
tail -n 12 gpurun_out/r2b_fus.log
timeout 300 python bench.py --workload fusion > gpurun_out/r2b_fusion.json 2> gpurun_out/r2b_fusion.err; echo "bench fusion rc=$?"
timeout 300 python bench.py --workload ctc --no-comparators > gpurun_out/r2b_ctc.json 2> gpurun_out/r2b_ctc.err; echo "bench ctc rc=$?"
timeout 300 python bench.py --workload hot --no-comparators > gpurun_out/r2b_hot.json 2> gpurun_out/r2b_hot.err; echo "bench hot rc=$?"
tail -c 600 gpurun_out/r2b_fusion.err
