#!/bin/bash
# 8 GPUs: 2-rank NCCL parity test, DDP step profile (exposed all-reduce), bench at N=8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -x -s > gpurun_out/r2n8_ddp_test.log 2>&1; echo "ddp test rc=$?" | tee -a gpurun_out/r2n8_ddp_test.log
tail -n 4 gpurun_out/r2n8_ddp_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/profile_ddp.py > gpurun_out/r2n8_profile.txt 2> gpurun_out/r2n8_profile.err; echo "profile rc=$?"
head -n 14 gpurun_out/r2n8_profile.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n8_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','allreduce_bytes_per_step')}); print(d['e2e']); print(d['hot_path']); print({k:d['beam'][k] for k in ('utt_per_s','ms')})
PY
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 4 --no-comparators --no-cpu-baseline > gpurun_out/r2n8_bench1.json 2> gpurun_out/r2n8_bench1.err; echo "bench1 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2n8_bench1.json')); print('N=1 on the same box:', d['value'], d['ms_per_step'])"
