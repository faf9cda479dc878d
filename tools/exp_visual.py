"""Experiment: VisualEncoder step time under memory-format / grad-mode variants (B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
vis = pkg.VisualEncoder().to(dev).train()
for p in vis.parameters(): p.requires_grad = False
x = torch.rand(8, 1, 150, 96, 96, device=dev)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): y = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, y
def base():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return vis(x)
ms, y0 = t(base); print("baseline NCHW autocast bf16: %.2f ms" % ms, flush=True)
def nograd():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return vis(x)
ms, y1 = t(nograd); print("no_grad: %.2f ms  maxdiff %.3g" % (ms, (y1.float() - y0.float()).abs().max().item()), flush=True)
vis2 = pkg.VisualEncoder().to(dev).train(); vis2.load_state_dict(vis.state_dict())
vis2.trunk.to(memory_format=torch.channels_last)
def cl():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        b = x.shape[0]
        y = vis2.frontend3D(x)
        tt, h, w = y.shape[2:]
        y = y.transpose(1, 2).reshape(b * tt, 64, h, w).contiguous(memory_format=torch.channels_last)
        return vis2.trunk(y).view(b, tt, 512)
ms, y2 = t(cl); print("trunk channels_last: %.2f ms  maxdiff %.3g" % (ms, (y2.float() - y0.float()).abs().max().item()), flush=True)
vis2.frontend3D.to(memory_format=torch.channels_last_3d)
def cl3():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        b = x.shape[0]
        y = vis2.frontend3D(x.contiguous(memory_format=torch.channels_last_3d))
        tt, h, w = y.shape[2:]
        y = y.transpose(1, 2).reshape(b * tt, 64, h, w).contiguous(memory_format=torch.channels_last)
        return vis2.trunk(y).view(b, tt, 512)
ms, y3 = t(cl3); print("trunk channels_last + frontend channels_last_3d: %.2f ms  maxdiff %.3g" % (ms, (y3.float() - y0.float()).abs().max().item()), flush=True)
# frontend only timings
def fe(): 
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return vis.frontend3D(x)
ms, _ = t(fe); print("frontend3D only (NCDHW): %.2f ms" % ms, flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    cl(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=90))
