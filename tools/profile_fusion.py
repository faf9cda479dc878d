import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
torch.manual_seed(0)
fus = pkg.CrossAttentionFusion(512, 1024, 512).to(dev)
B, Tv, Ta = 32, 150, 249
vis = torch.randn(B, Tv, 512, device=dev, dtype=torch.bfloat16)
aud = torch.randn(B, Ta, 1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
mask = torch.zeros(B, Ta, dtype=torch.long, device=dev); mask[:, :150] = 1; mask[:, 150:200] = 2
for b in range(B): mask[b, Ta - (b % 7):] = 3
r = torch.randn(B, Tv, 512, device=dev)
def fwd_bwd():
    fus.zero_grad(set_to_none=True); aud.grad = None
    f, _, _ = fus.fused_projection(vis, aud, mask)
    f.backward(r)
for _ in range(3): fwd_bwd()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): fwd_bwd()
torch.cuda.synchronize()
print("wall ms/iter", (time.perf_counter() - t0) / 10 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): fwd_bwd()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
# kernel-by-kernel list of one iteration
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
t00 = evs[2 * n].time_range.start
for e in evs[2 * n:]:
    print(f"{(e.time_range.start - t00):9.1f} us  dur {e.time_range.elapsed_us():7.1f}  {e.name[:80]}")
