"""BiLSTM fwd+bwd with the sentinel exchange (lstm_tag=1) vs the counter barrier (0): time and equality."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200 import _lib
import bench
dev = torch.device("cuda:0")
torch.manual_seed(0)
fus = pkg.CrossAttentionFusion(512, 1024, 512).to(dev)
flush = bench.l2_flusher(dev)
for B in (8, 16):
    x = torch.randn(B, 150, 512, device=dev, requires_grad=True)
    r = torch.randn(B, 150, 1024, device=dev)
    outs = {}
    for tag in (0, 3, 0, 3):
        _lib.set_tuning("lstm_tag", tag)
        def fb():
            x.grad = None; fus.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = fus.temporal(x)
            y.float().backward(r)
            return y
        y = fb()
        outs[tag] = (y.detach().float().clone(), x.grad.clone(), fus.temporal_model.weight_hh_l0.grad.clone())
        t, _ = bench.event_time(fb, 10, 3, flush, dev)
        def f():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                fus.temporal(x)
        tf, _ = bench.event_time(f, 10, 3, flush, dev)
        print(f"B={B} lstm_tag={tag}: fwd {tf*1e3:.0f} us  fwd+bwd {t*1e3:.0f} us", flush=True)
    same = all(torch.equal(a, b) for a, b in zip(outs[0], outs[3]))
    print("bitwise equal:", same, [float((a - b).abs().max()) for a, b in zip(outs[0], outs[3])])
