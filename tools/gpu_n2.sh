#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','e2e','gpu_launches','allreduce_bytes_per_step')}); print({k:d['beam'][k] for k in ('ms','utt_per_s','e2e_utt_per_s')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json
