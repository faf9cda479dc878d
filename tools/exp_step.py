"""Train-step variants: enqueue order x sync-free audio forward, resident and e2e; host enqueue time vs wall."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_av_model_b200.synthetic import make_batch
dev = torch.device("cuda:0")
tr = bench.build_models(dev)
host = make_batch(pairs=8, seconds=5.0, t_v=150, vocab=bench.VOCAB, seed=1234, pin=True)
devb = {k: v.to(dev) for k, v in host.items()}
def run(batch, read, n=8, warm=3):
    for _ in range(warm):
        tr.train_step(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); enq = 0.0
    for _ in range(n):
        a = time.perf_counter()
        l = tr.train_step(batch)
        enq += time.perf_counter() - a
        if read:
            float(l)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, enq / n * 1e3
for heavy in (False, True):
    for sf in (False, True):
        tr.gpu_heavy_first = heavy
        tr.audio_encoder.sync_free = sf
        w1, e1 = run(devb, False)
        w2, e2 = run(host, True)
        print(f"heavy_first={heavy} sync_free={sf}: resident {w1:.1f} ms (host enqueue {e1:.1f}) | e2e {w2:.1f} ms (host enqueue {e2:.1f})", flush=True)
# GPU busy time of one step (profiler), best variant
from torch.profiler import profile, ProfilerActivity
tr.gpu_heavy_first = True; tr.audio_encoder.sync_free = True
import sys; sys.exit(0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        float(tr.train_step(host))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
