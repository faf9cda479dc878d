"""Robustness probe of the probability-domain CTC kernels on very peaked log-probs (large logit scales)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
import multimodal_av_model_b200 as pkg
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_ctc_gpu import make_case, rel
for scale in (5.0, 20.0, 40.0, 80.0, 150.0):
    for lin in (1, 0):
        pkg._lib.set_tuning("ctc_lin", lin)
        lp, tg, il, tl = make_case(200, 8, 800, 3, 20, 58, seed=int(scale), scale=scale)
        ref = oracle.ctc_loss(lp.numpy(), tg, il, tl, blank=3, reduction="mean", zero_infinity=True)
        x = lp.cuda().requires_grad_()
        loss = pkg.ctc_loss(x, torch.from_numpy(tg).cuda(), torch.from_numpy(il).cuda(), torch.from_numpy(tl).cuda(), blank=3, reduction="mean", zero_infinity=True)
        loss.backward()
        g = x.grad.cpu().numpy()
        print(f"scale {scale:6.1f} lin={lin}: min lp {lp.min().item():9.1f}  loss {loss.item():12.4f} ref {ref['loss']:12.4f} relerr {abs(loss.item()-ref['loss'])/abs(ref['loss']):.2e}  grad relerr {rel(g, ref['grad']):.2e} finite {np.isfinite(g).all()}", flush=True)
pkg._lib.set_tuning("ctc_lin", 1)
