"""CTCDecoder — drop-in for /root/reference/model/decoder.py:6-35 on sm_100a kernels.

    CTCDecoder(input_dim, vocab_size, blank_id=0)
    forward(x[B,T,D], target=None, input_lengths=None, target_lengths=None) -> log_probs[B,T,V] | CTC loss

state_dict keys `net.0.weight`, `net.0.bias` as in the reference.  Linear + log_softmax are ONE tcgen05 cluster kernel
(csrc/ctc_head.cu: bf16 operands, fp32 accumulate in tensor memory, log_softmax in fp32 — what autocast does in the
reference — with the row statistics exchanged through distributed shared memory; the logits never reach HBM), the
optional loss branch (decoder.py:27-33) runs on the CTC kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .ctc import CTCLoss
from .gemm import gemm, operand

_BF16 = torch.bfloat16


class _CTCHeadFn(torch.autograd.Function):
    """log_softmax(x W^T + b) [passes = 2: log_softmax applied twice, as evaluate() does] with its backward.
    V <= 1024: ONE cluster kernel forward (logits stay in tensor memory, csrc/ctc_head.cu) and three launches backward
    (log_softmax backward; dgrad + wgrad grouped; bias column sum).  Larger vocabularies: GEMM + log_softmax kernels."""

    @staticmethod
    def forward(ctx, x, w, b, passes):
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
        bf = b.detach().float().contiguous()
        M = xb.shape[0]
        L = _lib.lib()
        lp = torch.empty((M, V), dtype=torch.float32, device=dev)
        fused = V <= 1024
        with _lib.device_guard(dev):
            if fused:
                _lib.check(L.avctc_ctc_head_forward(xb.data_ptr(), wb.data_ptr(), bf.data_ptr(), M, V, D, lp.data_ptr(),
                                                    int(passes), _lib.stream_ptr(dev)), "avctc_ctc_head_forward")
            else:
                logits = torch.empty((M, V), dtype=torch.float32, device=dev)
                gemm(operand(xb), operand(wb), M, V, D, logits, bias=bf, bias_mode=1)
                for _ in range(int(passes)):
                    _lib.check(L.avctc_log_softmax_forward(logits.data_ptr(), _lib.F32, lp.data_ptr(), _lib.F32, M, V,
                                                           _lib.stream_ptr(dev)), "avctc_log_softmax_forward")
                    logits = lp
        ctx.save_for_backward(xb, wb, lp)
        ctx.shp = (shp, x.dtype, fused, int(passes))
        return lp.view(*shp[:-1], V)

    @staticmethod
    def backward(ctx, dlp):
        xb, wb, lp = ctx.saved_tensors
        shp, xdtype, fused, passes = ctx.shp
        if passes != 1:
            raise RuntimeError("the doubly normalised head (passes=2) is the evaluation path: no backward")
        dev = dlp.device
        M, D = xb.shape
        V = wb.shape[0]
        Vp = (V + 7) // 8 * 8
        dy = dlp.reshape(M, V).float().contiguous()
        g_w = torch.empty((V, D), dtype=torch.float32, device=dev)
        g_b = torch.empty(V, dtype=torch.float32, device=dev)
        dz = torch.empty((M, Vp), dtype=_BF16, device=dev)
        dxb = torch.empty((M, D), dtype=_BF16, device=dev) if ctx.needs_input_grad[0] else None
        with _lib.device_guard(dev):
            _lib.check(_lib.lib().avctc_ctc_head_backward(lp.data_ptr(), dy.data_ptr(), xb.data_ptr(), wb.data_ptr(), M, V, D,
                                                          dz.data_ptr(), g_w.data_ptr(), g_b.data_ptr(),
                                                          dxb.data_ptr() if dxb is not None else None,
                                                          _lib.stream_ptr(dev)), "avctc_ctc_head_backward")
        dx = dxb.view(shp).to(xdtype) if dxb is not None else None
        return dx, g_w, g_b, None


class CTCDecoder(nn.Module):
    def __init__(self, input_dim, vocab_size, blank_id=0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, vocab_size))
        self.ctc_loss = CTCLoss(blank=blank_id, zero_infinity=True)

    def log_probs(self, x, passes=1):
        """[B,T,V] fp32 log-probs; passes=2 = F.log_softmax applied once more to them in the same kernel (what
        MultimodalTrainer.evaluate does with the decoder's output, trainer.py:212,221)."""
        lin = self.net[0]
        return _CTCHeadFn.apply(x, lin.weight, lin.bias, passes)

    def forward(self, x, target=None, input_lengths=None, target_lengths=None):
        log_probs = self.log_probs(x)                                   # [B, T, V] fp32
        if target is not None:
            return self.ctc_loss(log_probs.transpose(0, 1), target, input_lengths, target_lengths)
        return log_probs
