import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "hot"
tr = bench.build_models(dev, encoders=(what != "hot"))
from multimodal_av_model_b200.synthetic import make_features, make_batch
f = make_features(pairs=8, t_v=150, t_enc=249, seed=1234, dtype=torch.bfloat16)
fd = {k: [t.to(dev) for t in v] for k, v in f.items()}
for k in ("audio", "middle"):
    fd[k] = [t.requires_grad_() for t in fd[k]]
def hot_step():
    tr.optimizer.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        total = tr.hot_path_loss(fd["visual"], fd["audio"], fd["middle"], fd["masks"], fd["texts"], fd["lens"])[0]
    total.backward()
batch = {k: v.to(dev) for k, v in make_batch(pairs=8, seconds=5.0, t_v=150, seed=1234).items()}
fn = hot_step if what == "hot" else (lambda: tr.train_step(batch))
for _ in range(3): fn()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): fn()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): fn()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
# kernel-by-kernel timeline of the last iteration
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
t00 = evs[2 * n].time_range.start
busy = 0.0
for e in evs[2 * n:]:
    busy += e.time_range.elapsed_us()
    print(f"{(e.time_range.start - t00):9.1f} us  dur {e.time_range.elapsed_us():7.1f}  {e.name[:90]}")
print("GPU busy us in the last iteration:", busy, "span", evs[-1].time_range.end - t00)
