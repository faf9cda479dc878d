#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -n 3 gpurun_out/bench2.err
python tools/profile_hot.py full > gpurun_out/prof_full2.log 2>&1
head -c 600 gpurun_out/bench2.json; echo; grep -o '"hot_path": {[^}]*}' gpurun_out/bench2.json; grep -o '"beam": {[^}]*}' gpurun_out/bench2.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench2.json;  grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench2.json
