"""One CTC forward+backward at config 2 (B=64, T=1000, V=801) through the product's autograd route — the process ncu wraps
(`ncu --set full -k regex:ctc_ ...`; tools/ncu_summary.py traffic reads the LAST launch of each kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multimodal_av_model_b200 as pkg
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
lp, tg, il, tl, Lm = bench.ctc_case(T, dev)
for _ in range(2):
    x = lp.clone().requires_grad_()
    loss = pkg.ctc_loss(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
torch.cuda.synchronize()
print(float(loss), float(x.grad.abs().sum()))
