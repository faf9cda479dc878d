import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
vis = pkg.VisualEncoder().to(dev).train()
for p in vis.parameters(): p.requires_grad = False
x = torch.rand(8, 1, 150, 96, 96, device=dev)
def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return vis(x)
for _ in range(3): run()
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): y = run()
b.record(); torch.cuda.synchronize()
print("visual encoder fwd (2-D frontend): %.2f ms" % (a.elapsed_time(b) / 5), y.shape, y.dtype)
with torch.autocast("cuda", dtype=torch.bfloat16):
    y3 = vis.frontend3D(x); t, h, w = y3.shape[2:]
    ref = vis.trunk(y3.transpose(1, 2).reshape(8 * t, 64, h, w)).view(8, t, 512)
print("vs 3-D route: max diff %.4f (scale %.3f)" % ((ref.float() - y.float()).abs().max().item(), ref.float().abs().max().item()))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=80))
