#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "rc=$?"; tail -n 3 gpurun_out/bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','allreduce_bytes_per_step')}); print(d['e2e']); print({k:d['beam'][k] for k in ('ms','utt_per_s','e2e_utt_per_s')})
"
