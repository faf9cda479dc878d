"""GPU: the fused tcgen05 attention kernels (csrc/attention.cu) through the C ABI against the formulation the reference
runs inside nn.MultiheadAttention (fusion_module.py:61 -> torch/nn/functional.py:6630-6652): softmax(q k^T / sqrt(hd)) v
over all T keys, no mask — computed here in fp32 by torch on the SAME bf16 inputs, so the only differences are the
bf16 rounding of P / dS / outputs inside the kernel (tolerance: bf16, measured values printed)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("B,T,H", [(2, 150, 4), (1, 5, 2), (3, 37, 4), (2, 128, 1), (2, 130, 4), (1, 192, 2), (8, 64, 4),
                                   (32, 150, 4)])
def test_attention_forward_backward_vs_torch(B, T, H):
    pkg = _pkg()
    L = pkg._lib.lib()
    E = 128 * H
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    q = torch.randn(B, T, E, generator=g).to(dev).to(torch.bfloat16)
    kv = torch.randn(B, T, 2 * E, generator=g).to(dev).to(torch.bfloat16)
    dout = torch.randn(B, T, E, generator=g).to(dev).to(torch.bfloat16)
    # reference: fp32 math on the same bf16 values
    qf = q.float().view(B, T, H, 128).transpose(1, 2).requires_grad_()
    kf = kv[..., :E].float().view(B, T, H, 128).transpose(1, 2).requires_grad_()
    vf = kv[..., E:].float().view(B, T, H, 128).transpose(1, 2).requires_grad_()
    s = (qf @ kf.transpose(-1, -2)) / math.sqrt(128.0)
    o_ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, E)
    o_ref.backward(dout.float())
    lse_ref = torch.logsumexp(s, -1) / math.log(2.0)
    dq_ref = qf.grad.transpose(1, 2).reshape(B, T, E)
    dk_ref = kf.grad.transpose(1, 2).reshape(B, T, E)
    dv_ref = vf.grad.transpose(1, 2).reshape(B, T, E)
    o = torch.full((B, T, E), float("nan"), dtype=torch.bfloat16, device=dev)
    lse2 = torch.empty(B * H, T, dtype=torch.float32, device=dev)
    st = pkg._lib.stream_ptr(dev)
    pkg._lib.check(L.avctc_attention_forward(q.data_ptr(), kv.data_ptr(), o.data_ptr(), lse2.data_ptr(), B, T, H, E, st), "fwd")
    dq = torch.full((B, T, E), float("nan"), dtype=torch.bfloat16, device=dev)
    dkv = torch.full((B, T, 2 * E), float("nan"), dtype=torch.bfloat16, device=dev)
    pkg._lib.check(L.avctc_attention_backward(q.data_ptr(), kv.data_ptr(), dout.data_ptr(), o.data_ptr(), lse2.data_ptr(),
                                              dq.data_ptr(), dkv.data_ptr(), B, T, H, E, st), "bwd")
    torch.cuda.synchronize()
    errs = dict(o=relerr(o, o_ref), lse=(lse2.view(B, H, T) - lse_ref).abs().max().item(), dq=relerr(dq, dq_ref),
                dk=relerr(dkv[..., :E], dk_ref), dv=relerr(dkv[..., E:], dv_ref))
    print((B, T, H), {k: round(v, 5) for k, v in errs.items()})
    assert torch.isfinite(o.float()).all() and torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    assert errs["lse"] < 2e-3
    for k in ("o", "dq", "dk", "dv"):
        assert errs[k] < 2e-2, (k, errs)


def test_attention_rejects_unsupported_shapes():
    pkg = _pkg()
    L = pkg._lib.lib()
    dev = torch.device("cuda")
    x = torch.zeros(1, 200, 512, dtype=torch.bfloat16, device=dev)
    kv = torch.zeros(1, 200, 1024, dtype=torch.bfloat16, device=dev)
    l = torch.zeros(4, 200, device=dev)
    st = pkg._lib.stream_ptr(dev)
    assert L.avctc_attention_forward(x.data_ptr(), kv.data_ptr(), x.data_ptr(), l.data_ptr(), 1, 200, 4, 512, st) == -2   # T > 192
    assert L.avctc_attention_forward(x.data_ptr(), kv.data_ptr(), x.data_ptr(), l.data_ptr(), 1, 100, 8, 512, st) == -2   # hd != 128
