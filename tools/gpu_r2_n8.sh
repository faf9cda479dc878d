#!/bin/bash
# 8 GPUs: DDP step profile (exposed all-reduce) for bucket sizes / NCCL protocols, then bench at N=8 and N=1 on the same box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
AVCTC_BUCKETS=24,48,12 timeout 600 $TR --master-port 29555 tools/profile_ddp.py > gpurun_out/r2n8_profile.txt 2> gpurun_out/r2n8_profile.err; echo "profile rc=$?"
NCCL_PROTO=Simple AVCTC_NO_PROFILE=1 AVCTC_BUCKETS=24,48 timeout 600 $TR --master-port 29557 tools/profile_ddp.py >> gpurun_out/r2n8_profile.txt 2>> gpurun_out/r2n8_profile.err; echo "profile simple rc=$?"
grep "^world" gpurun_out/r2n8_profile.txt
timeout 900 $TR --master-port 29556 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n8_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','allreduce_bytes_per_step')}); print(d['e2e']); print(d['hot_path']); print({k:d['beam'][k] for k in ('utt_per_s','ms')})
PY
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 4 --no-comparators --no-cpu-baseline > gpurun_out/r2n8_bench1.json 2> gpurun_out/r2n8_bench1.err; echo "bench1 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2n8_bench1.json')); print('N=1 on the same box:', d['value'], d['ms_per_step'])"
