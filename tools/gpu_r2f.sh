#!/bin/bash
# round 2 evidence pass: full GPU suite, the driver's bench command (ours + reference arm), knob matrix, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2f_all.log
tail -n 4 gpurun_out/r2f_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2f_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 1200 python tools/knob_matrix.py > gpurun_out/r2f_knob_matrix.txt 2>&1; tail -n 1 gpurun_out/r2f_knob_matrix.txt
python tools/run_ctc_once.py > gpurun_out/r2f_ctc_once.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_ -o gpurun_out/r2f_ctc -f python tools/run_ctc_once.py > gpurun_out/r2f_ncu_ctc.log 2>&1; echo "ncu ctc rc=$?"
python tools/run_fusion_once.py > gpurun_out/r2f_fusion_once.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attention_kernel|ctc_head_fwd|gemm_bf16" --launch-skip 12 -c 24 -o gpurun_out/r2f_fusion -f python tools/run_fusion_once.py > gpurun_out/r2f_ncu_fusion.log 2>&1; echo "ncu fusion rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_hot_launches.csv python bench.py --workload hot --no-comparators --steps 2 --warmup 3 > gpurun_out/r2f_ncu_hot.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_hot.py > gpurun_out/r2f_hot_timeline.txt 2>&1; grep -n "wall ms\|GPU busy" gpurun_out/r2f_hot_timeline.txt
