#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_gpu.py tests/test_bench_sizes_gpu.py tests/test_fusion_gpu.py tests/test_trainer_gpu.py -q -x > gpurun_out/r2g_lstm.log 2>&1; echo "lstm tests rc=$?" | tee -a gpurun_out/r2g_lstm.log
tail -n 8 gpurun_out/r2g_lstm.log
timeout 300 python bench.py --workload lstm > gpurun_out/r2g_lstm.json 2> gpurun_out/r2g_lstm.err; echo "bench lstm rc=$?"
python -c "import json; print(json.load(open('gpurun_out/r2g_lstm.json'))['lstm'])"
timeout 300 python bench.py --workload hot --no-comparators > gpurun_out/r2g_hot.json 2> gpurun_out/r2g_hot.err; echo "bench hot rc=$?"
python -c "import json; print(json.load(open('gpurun_out/r2g_hot.json'))['hot_path'])"
