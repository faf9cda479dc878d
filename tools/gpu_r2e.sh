#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2e_all.log
tail -n 6 gpurun_out/r2e_all.log
timeout 300 python bench.py --workload hot --no-comparators > gpurun_out/r2e_hot.json 2> gpurun_out/r2e_hot.err; echo "bench hot rc=$?"
python -c "import json; print(json.load(open('gpurun_out/r2e_hot.json'))['hot_path'])"
python tools/profile_hot.py > gpurun_out/r2e_hot_timeline.txt 2>&1; echo "profile rc=$?"; grep -n "wall ms\|GPU busy" gpurun_out/r2e_hot_timeline.txt
