mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ctc_gpu.py tests/test_bench_sizes_gpu.py -q -x 2>&1 | tail -2
timeout 250 python tools/exp_ctc_overlap.py 2>&1 | tail -10 | cut -c1-150
python tools/run_ctc_once.py > gpurun_out/r2f_ctc_once.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_ -o gpurun_out/r2f_ctc -f python tools/run_ctc_once.py > gpurun_out/r2f_ncu_ctc.log 2>&1; echo "ncu ctc rc=$?"
