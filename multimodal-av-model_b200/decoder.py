"""CTCDecoder — drop-in for /root/reference/model/decoder.py:6-35 on sm_100a kernels.

    CTCDecoder(input_dim, vocab_size, blank_id=0)
    forward(x[B,T,D], target=None, input_lengths=None, target_lengths=None) -> log_probs[B,T,V] | CTC loss

state_dict keys `net.0.weight`, `net.0.bias` as in the reference.  The linear runs on the tcgen05 GEMM
(bf16 operands, fp32 accumulate), log_softmax in fp32 (what autocast does in the reference), the optional
loss branch (decoder.py:27-33) on the CTC kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .ctc import CTCLoss
from .gemm import gemm, operand

_BF16 = torch.bfloat16


class _CTCHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        _lib.require_cuda(x, "x")
        dev = x.device
        shp = x.shape
        D = shp[-1]
        V = w.shape[0]
        if D % 8:
            raise RuntimeError("input_dim must be a multiple of 8 (TMA row alignment)")
        xb = x.detach().reshape(-1, D)
        xb = (xb if xb.dtype == _BF16 else xb.to(_BF16)).contiguous()
        wb = w.detach().to(_BF16).contiguous()
        M = xb.shape[0]
        logits = torch.empty((M, V), dtype=torch.float32, device=dev)
        gemm(operand(xb), operand(wb), M, V, D, logits, bias=b.detach().float().contiguous(), bias_mode=1)
        lp = torch.empty((M, V), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().avctc_log_softmax_forward(logits.data_ptr(), _lib.F32, lp.data_ptr(), _lib.F32, M, V,
                                                            _lib.stream_ptr(dev)), "avctc_log_softmax_forward")
        ctx.save_for_backward(xb, wb, lp)
        ctx.shp = (shp, x.dtype)
        return lp.view(*shp[:-1], V)

    @staticmethod
    def backward(ctx, dlp):
        xb, wb, lp = ctx.saved_tensors
        shp, xdtype = ctx.shp
        dev = dlp.device
        M, D = xb.shape
        V = wb.shape[0]
        Vp = (V + 7) // 8 * 8
        dy = dlp.reshape(M, V).float().contiguous()
        dz = torch.zeros((M, Vp), dtype=_BF16, device=dev) if Vp != V else torch.empty((M, Vp), dtype=_BF16, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().avctc_log_softmax_backward(lp.data_ptr(), dy.data_ptr(), _lib.F32, dz.data_ptr(), M, V, Vp,
                                                             _lib.stream_ptr(dev)), "avctc_log_softmax_backward")
        g_w = torch.empty((V, D), dtype=torch.float32, device=dev)
        gemm(operand(dz, "mn", rows=V), operand(xb, "mn"), V, D, M, g_w)
        g_b = torch.empty(V, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().avctc_colsum(dz.data_ptr(), _lib.BF16, M, V, Vp, g_b.data_ptr(), 0, _lib.stream_ptr(dev)),
                       "avctc_colsum")
        dx = None
        if ctx.needs_input_grad[0]:
            dxb = torch.empty((M, D), dtype=_BF16, device=dev)
            gemm(operand(dz, kdim=V), operand(wb, "mn"), M, D, V, dxb)
            dx = dxb.view(shp).to(xdtype)
        return dx, g_w, g_b


class CTCDecoder(nn.Module):
    def __init__(self, input_dim, vocab_size, blank_id=0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, vocab_size))
        self.ctc_loss = CTCLoss(blank=blank_id, zero_infinity=True)

    def forward(self, x, target=None, input_lengths=None, target_lengths=None):
        lin = self.net[0]
        log_probs = _CTCHeadFn.apply(x, lin.weight, lin.bias)          # [B, T, V] fp32
        if target is not None:
            return self.ctc_loss(log_probs.transpose(0, 1), target, input_lengths, target_lengths)
        return log_probs
