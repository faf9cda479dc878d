import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from conftest import load_cases
from oracle import torch_port as tp
import multimodal_av_model_b200 as pkg
def relerr(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
FUS = load_cases("fusion_cases.npz")
for name, c in FUS.items():
    dv, da = c["visual"].shape[-1], c["audio"].shape[-1]
    e = c["param/fusion_proj.weight"].shape[0]; H = int(c["num_heads"])
    sd = {k[6:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("param/")}
    fus = pkg.CrossAttentionFusion(dv, da, e, num_heads=H); fus.load_state_dict(sd); fus.cuda()
    ref = tp.FusionPort(dv, da, e, num_heads=H); ref.load_state_dict(sd)
    refb = tp.FusionPort(dv, da, e, num_heads=H); refb.load_state_dict(sd); refb.cuda()
    mask = torch.from_numpy(c["mask"])
    r = torch.randn(c["visual"].shape[0], c["visual"].shape[1], e)
    outs = {}
    for tag, mod, dev, ac in (("ref32", ref, "cpu", False), ("torch_bf16", refb, "cuda", True), ("ours", fus, "cuda", False)):
        vis = torch.from_numpy(c["visual"]).to(dev).requires_grad_(); aud = torch.from_numpy(c["audio"]).to(dev).requires_grad_()
        mod.zero_grad()
        if tag == "ours":
            f, m, il = mod.fused_projection(vis, aud, mask.to(dev))
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                f, m = mod.projection(vis, aud, mask.to(dev))
        (f.float() * r.to(dev)).sum().backward()
        outs[tag] = dict(f=f, dvis=vis.grad, daud=aud.grad, **{k: p.grad for k, p in mod.named_parameters() if p.grad is not None})
    print("==", name)
    for k in outs["ref32"]:
        print(f"  {k:40s} ours {relerr(outs['ours'][k], outs['ref32'][k]):.3e}   torch_bf16 {relerr(outs['torch_bf16'][k], outs['ref32'][k]):.3e}")
