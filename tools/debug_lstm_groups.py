import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200.fusion_module import _BiLSTMFn
for (B, T, H) in [(24, 20, 256), (24, 20, 512), (32, 20, 256), (20, 20, 256), (24, 3, 256), (17, 20, 256)]:
    torch.manual_seed(B * 7 + T)
    ref = torch.nn.LSTM(H, H, num_layers=2, batch_first=True, bidirectional=True).cuda()
    x = torch.randn(B, T, H, device="cuda")
    r = torch.randn(B, T, 2 * H, device="cuda")
    res = {}
    for groups in (1, 2, 1, 2):
        pkg._lib.set_tuning("lstm_groups", groups)
        xi = x.clone().requires_grad_()
        y = _BiLSTMFn.apply(xi, *ref._flat_weights)
        (y.float() * r).sum().backward()
        torch.cuda.synchronize()
        key = (groups, len([k for k in res if k[0] == groups]))
        res[key] = (y.detach().float().clone(), xi.grad.clone())
    pkg._lib.set_tuning("lstm_groups", 0)
    a, b = res[(1, 0)], res[(2, 0)]
    dy = (a[0] - b[0]).abs(); dx = (a[1] - b[1]).abs()
    rows = dy.amax(dim=(1, 2)).nonzero().flatten().tolist()
    ts = dy.amax(dim=(0, 2)).nonzero().flatten().tolist()
    print((B, T, H), "y diff", float(dy.max()), "rows", rows[:12], "frames", ts[:8], "| dx diff", float(dx.max()),
          "| repeat g1", float((res[(1, 0)][0] - res[(1, 1)][0]).abs().max()), "repeat g2", float((res[(2, 0)][0] - res[(2, 1)][0]).abs().max()),
          "cols(fwd half / bwd half)", float(dy[..., :H].max()), float(dy[..., H:].max()))
