"""Config-3 fusion forward+backward (B=32, T_v=150, T_a=249, bf16) and one CTC-head forward at the training size
(M = 16 x 150 rows, V=800, K=1024) — the process ncu wraps for the attention / head / grouped-GEMM captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
fus = pkg.CrossAttentionFusion(512, 1024, 512).to(dev)
dec = pkg.CTCDecoder(1024, 800, blank_id=3).to(dev)
B, Tv, Ta = 32, 150, 249
vis = torch.randn(B, Tv, 512, device=dev, dtype=torch.bfloat16)
aud = torch.randn(B, Ta, 1024, device=dev, dtype=torch.bfloat16, requires_grad=True)
mask = torch.zeros(B, Ta, dtype=torch.long, device=dev)
mask[:, :150] = 1; mask[:, 150:200] = 2
r = torch.randn(B, Tv, 512, device=dev, dtype=torch.bfloat16)
x = torch.randn(16, 150, 1024, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    fus.zero_grad(set_to_none=True); aud.grad = None
    f, _, _ = fus.fused_projection(vis, aud, mask)
    f.backward(r)
    with torch.no_grad():
        lp = dec(x)
torch.cuda.synchronize()
print(float(f.float().abs().mean()), float(lp.exp().sum(-1).mean()))
