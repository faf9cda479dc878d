"""Host wrapper of avctc_gemm_bf16 (tcgen05/TMEM/TMA): C = alpha * A . B^T (+ bias) on bf16 operands.

Internal to the fusion path (fusion_module.py, decoder.py); not part of the reference's call surface.
An operand is (tensor, 'k' | 'mn'): 'k' means the 2-D/3-D tensor is [.., rows, K] (row-major, K contiguous),
'mn' means it is [.., K, rows] (the transposed view is what enters the product; no copy is made).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def operand(t: torch.Tensor, major: str = "k", *, k_outer=0, k_inner=0, r_outer=0, r_inner=0, z_outer=0, z_inner=0,
            rows=None, kdim=None):
    """Describe a bf16 tensor ([R,C] or [Z,R,C], last dim contiguous) as a GEMM operand."""
    if t.dtype != torch.bfloat16:
        raise RuntimeError("gemm operands must be bfloat16")
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.stride(2) != 1:
        raise RuntimeError("gemm operand must be [R,C] or [Z,R,C] with a contiguous last dim")
    Z, R, Cc = t.shape
    op = _lib.GemmOperand()
    op.ptr = t.data_ptr()
    mn = major == "mn"
    op.rows = (Cc if mn else R) if rows is None else rows
    op.kdim = (R if mn else Cc) if kdim is None else kdim
    op.zdim = Z
    op.ld = t.stride(1)
    op.zstride = t.stride(0) if Z > 1 else 0
    op.k_outer, op.k_inner, op.r_outer, op.r_inner = k_outer, k_inner, r_outer, r_inner
    op.z_outer, op.z_inner = z_outer, z_inner
    op.mn_major = 1 if mn else 0
    return op, t


def gemm(a, b, M, N, K, out: torch.Tensor, *, batch=1, inner_count=1, ldc=None, c_outer=0, c_inner=0, bias=None,
         bias_mode=0, alpha=1.0, accumulate=False):
    """Launch on the current stream.  a, b: results of operand().  out: fp32 or bf16 tensor (written in place)."""
    (opa, ta), (opb, tb) = a, b
    _lib.require_cuda(out, "out")
    if ldc is None:
        ldc = out.stride(-2)
    if bias is not None and bias.dtype != torch.float32:
        raise RuntimeError("bias must be float32")
    dev = out.device
    with _lib.device_guard(dev):
        _lib.check(_lib.lib().avctc_gemm_bf16(
            ctypes.byref(opa), ctypes.byref(opb), int(M), int(N), int(K), int(batch), int(inner_count),
            out.data_ptr(), _lib.dtype_enum(out), int(ldc), int(c_outer), int(c_inner),
            bias.data_ptr() if bias is not None else None, int(bias_mode if bias is not None else 0),
            float(alpha), int(accumulate), _lib.stream_ptr(dev)), "avctc_gemm_bf16")
    return out


def linear_nt(x: torch.Tensor, w: torch.Tensor, bias=None, out_dtype=torch.bfloat16, alpha=1.0):
    """y[M,N] = x[M,K] . w[N,K]^T + bias  (nn.Linear forward)."""
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    return gemm(operand(x), operand(w), M, N, K, out, bias=bias, bias_mode=1, alpha=alpha)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM with its backward (dgrad, wgrad, bias colsum); bf16 operands."""

    @staticmethod
    def forward(ctx, x, w, b):
        shp = x.shape
        K = shp[-1]
        xb = x.detach().reshape(-1, K)
        xb = (xb if xb.dtype == torch.bfloat16 else xb.to(torch.bfloat16)).contiguous()
        wb = w.detach().to(torch.bfloat16).contiguous()
        y = linear_nt(xb, wb, b.detach().float().contiguous() if b is not None else None, out_dtype=torch.bfloat16)
        ctx.save_for_backward(xb, wb)
        ctx.meta = (shp, x.dtype, b is not None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        shp, xdtype, has_b = ctx.meta
        N, K = wb.shape
        M = xb.shape[0]
        d = dy.reshape(M, N)
        d = (d if d.dtype == torch.bfloat16 else d.to(torch.bfloat16)).contiguous()
        dev = d.device
        g_w = torch.empty((N, K), dtype=torch.float32, device=dev)
        gemm(operand(d, "mn"), operand(xb, "mn"), N, K, M, g_w)
        g_b = None
        if has_b:
            g_b = torch.empty(N, dtype=torch.float32, device=dev)
            with _lib.device_guard(dev):
                _lib.check(_lib.lib().avctc_colsum(d.data_ptr(), _lib.BF16, M, N, N, g_b.data_ptr(), 0,
                                                   _lib.stream_ptr(dev)), "avctc_colsum")
        dx = None
        if ctx.needs_input_grad[0]:
            dxb = torch.empty((M, K), dtype=torch.bfloat16, device=dev)
            gemm(operand(d), operand(wb, "mn"), M, K, N, dxb)
            dx = dxb.view(shp).to(xdtype)
        return dx, g_w, g_b
