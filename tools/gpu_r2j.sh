#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ctc_gpu.py tests/test_bench_sizes_gpu.py -q -x > gpurun_out/r2j_ctc.log 2>&1; echo "ctc tests rc=$?" | tee -a gpurun_out/r2j_ctc.log
tail -n 5 gpurun_out/r2j_ctc.log
for st in 1 0; do
python - <<PY
import torch, sys
sys.path.insert(0, '.')
import bench
import multimodal_av_model_b200 as pkg
pkg._lib.set_tuning("ctc_stage", $st)
dev = torch.device("cuda:0")
r = bench.bench_ctc(dev, comparators=False)
for T in ("T250", "T1000"):
    print("ctc_stage=$st", T, {k: round(r[T][k], 4) for k in ("product_ms", "abi_fwd_bwd_ms", "scan_ms", "grad_ms", "gbs", "grad_kernel_gbs")})
PY
done
