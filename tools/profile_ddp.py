"""Under torchrun (N ranks): the full config-4 train step with the bucketed gradient all-reduce, instrumented.
Rank 0 prints (1) step time, (2) the part of the all-reduce that is EXPOSED after backward's last kernel (CUDA events
around GradBucketReducer.finish()), (3) a torch-profiler table of one step (NCCL kernels and the top compute kernels).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/profile_ddp.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from multimodal_av_model_b200 import ddp
from multimodal_av_model_b200.synthetic import make_batch

rank, local, world = ddp.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
tr = bench.build_models(dev)
batch = {k: v.to(dev) for k, v in make_batch(pairs=bench.PAIRS_PER_GPU, seconds=bench.SECONDS, t_v=bench.T_V, seed=1234 + rank).items()}
from multimodal_av_model_b200.ddp import GradBucketReducer
never = tr.fusion_module.never_used_parameters()


def measure(bucket_mb, n=10):
    """Step time (max over ranks) and the part of the reduction exposed behind backward, for one bucket size."""
    if tr._reducer is not None:
        tr._reducer.close()
    red = tr._reducer = GradBucketReducer(tr.parameters, bucket_bytes=bucket_mb << 20, never_used=never) if world > 1 else None
    marks = []
    if red is not None:
        orig = red.finish
        def finish(ok=True):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()                      # behind backward's last kernel on the main stream
            orig(ok)
            b.record()
            marks.append((a, b))
        red.finish = finish
    for _ in range(4):
        tr.train_step(batch)
    torch.cuda.synchronize(); marks.clear()
    if world > 1:
        dist.barrier()
    a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(n):
        tr.train_step(batch)
    b0.record()
    torch.cuda.synchronize()
    dt = a0.elapsed_time(b0) / n
    exposed = sum(a.elapsed_time(b) for a, b in marks) / max(len(marks), 1)
    t = torch.tensor([dt, exposed], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        nb = len(red.buckets) if red is not None else 0
        print(f"world {world} NCCL_PROTO={os.environ.get('NCCL_PROTO', 'default')} bucket {bucket_mb} MB: {float(t[0]):.2f} ms/step "
              f"(CUDA events, max over ranks); exposed behind backward (finish(): tail buckets + wait + scale) {float(t[1]):.3f} ms; "
              f"{nb} buckets, {red.grad_bytes() / 1e6 if red else 0:.1f} MB of gradients per step", flush=True)


for mb in [int(x) for x in os.environ.get("AVCTC_BUCKETS", "24").split(",")]:
    measure(mb)
if os.environ.get("AVCTC_NO_PROFILE"):
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    nccl = [e for e in evs if "nccl" in e.name.lower()]
    span0, span1 = min(e.time_range.start for e in evs), max(e.time_range.end for e in evs)
    print(f"profiled step: GPU span {(span1 - span0) / 1e3:.2f} ms, {len(evs)} kernels; NCCL kernels: {len(nccl)}, "
          f"{sum(e.time_range.elapsed_us() for e in nccl) / 1e3:.3f} ms total")
    for e in nccl:
        print(f"   nccl @ {(e.time_range.start - span0) / 1e3:8.3f} ms  dur {e.time_range.elapsed_us():8.1f} us  {e.name[:80]}")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
