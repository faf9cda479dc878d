#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_beam_gpu.py tests/test_lstm_gpu.py -q -x > gpurun_out/r2h_beam.log 2>&1; echo "beam+lstm tests rc=$?" | tee -a gpurun_out/r2h_beam.log
tail -n 8 gpurun_out/r2h_beam.log
for c in 1 0 4 8; do
python - <<PY
import torch, sys
sys.path.insert(0, '.')
import bench
import multimodal_av_model_b200 as pkg
pkg._lib.set_tuning("beam_chunks", $c)
dev = torch.device("cuda:0")
r = bench.bench_beam(dev, 0, 1, comparators=False)
print("beam_chunks=$c", {k: r[k] for k in ("ms", "utt_per_s_shard", "hbm_frac")})
PY
done
