"""Beam decode at config 5 (4096 x [150, 800], beam 10), once per kernel route — the process ncu wraps
(`ncu --set full -k regex:beam_ ...`): the two-phase kernels, then the fused kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200 import _lib
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(7)
lp = (3 * torch.randn(N, 150, 800, generator=g, device=dev)).log_softmax(-1)
res = []
for fused in (0, 1):
    _lib.set_tuning("beam_fused", fused)
    for _ in range(2):
        r = pkg.beam_search_batch(lp, beam_width=10, blank=3)
    res.append(r)
torch.cuda.synchronize()
print(len(res[0]), res[0] == res[1])
