"""GPU: tcgen05 GEMM (through the C ABI) against torch matmul on the same bf16 inputs.
Tolerance: inputs are identical bf16 values and both sides accumulate in fp32, so results agree to
fp32 summation-order noise (rel 1e-3 of the row scale covers K=1024); bf16 outputs add one rounding."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    import multimodal_av_model_b200 as pkg
    from multimodal_av_model_b200 import gemm
    return pkg, gemm


def check(out, ref, tol=2e-3):
    scale = ref.abs().max().item() + 1e-6
    err = (out.float() - ref).abs().max().item() / scale
    assert err < tol, err


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 512), (4800, 512, 1024), (150, 150, 128), (333, 800, 1024),
                                   (100, 72, 40)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_nt_kmajor(M, N, K, out_dtype):
    _, g = _mods()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = g.linear_nt(a, b, bias, out_dtype=out_dtype, alpha=0.5)
    ref = 0.5 * (a.float() @ b.float().t()) + bias
    check(out, ref, 2e-3 if out_dtype == torch.float32 else 8e-3)


@pytest.mark.parametrize("amn,bmn", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (512, 1024, 4800), (200, 136, 152)])
def test_gemm_mn_major(amn, bmn, M, N, K):
    _, g = _mods()
    torch.manual_seed(M * 3 + N + K + amn * 2 + bmn)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    ta = a.t().contiguous() if amn else a           # [K,M] storage when MN-major
    tb = b.t().contiguous() if bmn else b
    out = torch.empty(M, N, device="cuda")
    g.gemm(g.operand(ta, "mn" if amn else "k"), g.operand(tb, "mn" if bmn else "k"), M, N, K, out)
    check(out, a.float() @ b.float().t())
    out2 = torch.ones(M, N, device="cuda")
    g.gemm(g.operand(ta, "mn" if amn else "k"), g.operand(tb, "mn" if bmn else "k"), M, N, K, out2, accumulate=True)
    check(out2, a.float() @ b.float().t() + 1.0)


def test_gemm_batched_head_slices():
    """Attention-style addressing: Q,K are [B*T, E]; batch z=(b,h) reads rows b*T.., columns h*hd.. ."""
    _, g = _mods()
    torch.manual_seed(0)
    B, T, H, hd = 3, 150, 4, 128
    E = H * hd
    q = torch.randn(B * T, E, device="cuda").bfloat16()
    k = torch.randn(B * T, E, device="cuda").bfloat16()
    Tp = 152
    s = torch.zeros(B * H, T, Tp, device="cuda")
    g.gemm(g.operand(q, k_inner=hd, r_outer=T), g.operand(k, k_inner=hd, r_outer=T), T, T, hd, s,
           batch=B * H, inner_count=H, ldc=Tp, c_outer=H * T * Tp, c_inner=T * Tp, alpha=hd ** -0.5)
    qh = q.float().view(B, T, H, hd).permute(0, 2, 1, 3)
    kh = k.float().view(B, T, H, hd).permute(0, 2, 1, 3)
    ref = (qh @ kh.transpose(-1, -2)) * hd ** -0.5
    check(s[:, :, :T].view(B, H, T, T), ref)
    assert torch.all(s[:, :, T:] == 0)
    # P.V with V in its natural [B*T, E] layout as an MN-major B operand; P's padded tail must not leak
    p = torch.softmax(s[:, :, :T], -1)
    pb = torch.zeros(B * H, T, Tp, device="cuda", dtype=torch.bfloat16)
    pb[:, :, :T] = p.bfloat16()
    v = torch.randn(B * T, E, device="cuda").bfloat16()
    o = torch.empty(B * T, E, device="cuda", dtype=torch.bfloat16)
    g.gemm(g.operand(pb, z_outer=H, z_inner=1, kdim=T), g.operand(v, "mn", k_outer=T, r_inner=hd), T, hd, T, o,
           batch=B * H, inner_count=H, ldc=E, c_outer=T * E, c_inner=hd)
    vh = v.float().view(B, T, H, hd).permute(0, 2, 1, 3)
    ref_o = (pb[:, :, :T].float().view(B, H, T, T) @ vh).permute(0, 2, 1, 3).reshape(B * T, E)
    check(o, ref_o, 8e-3)
