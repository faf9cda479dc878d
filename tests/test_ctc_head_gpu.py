"""GPU: the fused CTC head (csrc/ctc_head.cu, cluster kernel) against log_softmax(x W^T + b) computed by torch in fp32 on
the same bf16-rounded operands (decoder.py:24-25), forward with one and two normalisation passes, and its backward."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


@pytest.mark.parametrize("M,V,K", [(2400, 800, 1024), (77, 801, 64), (130, 30, 64), (300, 1024, 128), (1, 800, 1024),
                                   (257, 129, 72), (200, 1100, 64)])
def test_head_forward_backward_vs_torch(M, V, K):
    pkg = _pkg()
    torch.manual_seed(M + V)
    dec = pkg.CTCDecoder(K, V, blank_id=3).cuda()
    x = torch.randn(3, (M + 2) // 3, K, device="cuda")[:, :, :].reshape(-1, K)[:M].reshape(1, M, K).contiguous().requires_grad_()
    lp = dec(x)
    w, b = dec.net[0].weight, dec.net[0].bias
    x2 = x.detach().clone().requires_grad_()
    w2 = w.detach().to(torch.bfloat16).float().requires_grad_()
    b2 = b.detach().clone().requires_grad_()
    ref = F.log_softmax(x2.to(torch.bfloat16).float() @ w2.t() + b2, dim=-1)
    r = torch.randn_like(ref)
    (lp * r).sum().backward()
    (ref * r).sum().backward()
    err = (lp - ref).abs().max().item()
    print((M, V, K), "max|dlp|", round(err, 6), "sum(exp) - 1:", round((lp.exp().sum(-1) - 1).abs().max().item(), 7))
    assert err < 2e-4                                      # same operands, fp32 accumulation: only summation order differs
    assert (lp.exp().sum(-1) - 1).abs().max().item() < 1e-4
    rel = lambda a, c: ((a - c).abs().max() / (c.abs().max() + 1e-12)).item()
    assert rel(b.grad, b2.grad) < 1e-2                    # dz is rounded to bf16 before the GEMMs
    assert rel(w.grad, w2.grad) < 1e-2
    assert rel(x.grad, x2.grad) < 1e-2


@pytest.mark.parametrize("M,V,K", [(2400, 800, 1024), (50, 801, 64)])
def test_head_two_passes_equals_log_softmax_twice(M, V, K):
    """evaluate() normalises the decoder's log-probs once more (trainer.py:212,221); passes=2 does it in the kernel."""
    pkg = _pkg()
    torch.manual_seed(7)
    dec = pkg.CTCDecoder(K, V, blank_id=3).cuda()
    x = torch.randn(2, M // 2, K, device="cuda")
    with torch.no_grad():
        once = dec(x)
        twice = dec.log_probs(x, passes=2)
        ref = F.log_softmax(once, dim=-1)
    assert (twice - ref).abs().max().item() < 2e-6
    assert (twice - once).abs().max().item() < 1e-5       # idempotent up to rounding
