"""One BiLSTM forward+backward (B=16, T=150, H=512: both speakers of a config-4 step) — the process ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
fus = pkg.CrossAttentionFusion(512, 1024, 512).to(dev)
x = torch.randn(16, 150, 512, device=dev, requires_grad=True)
r = torch.randn(16, 150, 1024, device=dev)
for _ in range(2):
    x.grad = None; fus.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = fus.temporal(x)
    y.float().backward(r)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
