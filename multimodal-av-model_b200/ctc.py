"""CTC loss with the call surface of torch.nn.CTCLoss, backed by the sm_100a kernels.

Drop-in for `nn.CTCLoss(blank=tokenizer.blank_id, zero_infinity=True)` as constructed at
/root/reference/model/trainer.py:25 (and model/decoder.py:12) and called at trainer.py:116-117,224-225
with `(log_probs[T,B,V] (transposed view of [B,T,V]), targets[B,L] int64, input_lengths[B], target_lengths[B])`.
Differences from torch that matter: lengths are read ON THE DEVICE (torch copies them to the host every
call), the strided [T,B,V] view is consumed without a copy, and bfloat16 log-probs are accepted
(torch has no half/bf16 CTC kernel).  Gradient convention is ATen's: exp(lp) - posterior.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


def _as_i64(t, device, name):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, dtype=torch.long)
    if t.dtype != torch.long:
        t = t.long()
    if t.device != device:
        t = t.to(device)
    return t.contiguous()


class _CTCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity):
        _lib.require_cuda(log_probs, "log_probs")
        if log_probs.dim() != 3:
            raise RuntimeError("log_probs must be (T, B, V)")
        if log_probs.stride(2) != 1:
            log_probs = log_probs.contiguous()
        dev = log_probs.device
        T, B, V = log_probs.shape
        targets = _as_i64(targets, dev, "targets")
        input_lengths = _as_i64(input_lengths, dev, "input_lengths")
        target_lengths = _as_i64(target_lengths, dev, "target_lengths")
        if input_lengths.numel() != B or target_lengths.numel() != B:
            raise RuntimeError("input_lengths and target_lengths must have batch_size elements")
        offsets = None
        if targets.dim() == 2:
            if targets.size(0) != B:
                raise RuntimeError("targets must be (B, L) or 1-D concatenated")
            lmax, tstride = int(targets.size(1)), int(targets.stride(0))
        elif targets.dim() == 1:   # concatenated targets (not used by the reference): one host sync
            offsets = (torch.cumsum(target_lengths, 0) - target_lengths).contiguous()
            lmax, tstride = (int(target_lengths.max().item()) if B else 0), 0
        else:
            raise RuntimeError("targets must be 1-D or 2-D")
        need_grad = bool(ctx.needs_input_grad[0])
        L = _lib.lib()
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ws_bytes = int(L.avctc_ctc_workspace_bytes(T, B, lmax)) if need_grad else 0
        if need_grad and ws_bytes == 0:
            raise RuntimeError("CTC: target length not supported by the sm_100a kernels")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if need_grad else None
        st = _lib.stream_ptr(dev)
        red = _lib.REDUCTION[reduction]
        grad = None
        out = torch.empty(B if red == 0 else 1, dtype=torch.float32, device=dev)
        off_ptr = offsets.data_ptr() if offsets is not None else None
        with _lib.device_guard(dev):
            if need_grad:
                # The gradient pass goes out NOW, directly behind the scan, with a unit grad_out: launched there it
                # starts on each utterance as soon as that utterance's alpha/beta rows are complete, under the scans of
                # the longer ones (csrc/ctc_loss.cu, "early" route).  backward() then only applies the incoming factor.
                # The lattice workspace dies with this call instead of living until backward.
                grad = torch.empty((T, B, V), dtype=log_probs.dtype, device=dev)
                _lib.check(L.avctc_ctc_forward_backward(
                    log_probs.data_ptr(), _lib.dtype_enum(log_probs), log_probs.stride(0), log_probs.stride(1),
                    T, B, V, targets.data_ptr(), tstride, off_ptr, input_lengths.data_ptr(), target_lengths.data_ptr(),
                    lmax, int(blank), red, int(zero_infinity), nll.data_ptr(), out.data_ptr(), _unit(dev).data_ptr(), 0,
                    grad.data_ptr(), ws.data_ptr(), ws_bytes, st), "avctc_ctc_forward_backward")
            else:
                _lib.check(L.avctc_ctc_forward(
                    log_probs.data_ptr(), _lib.dtype_enum(log_probs), log_probs.stride(0), log_probs.stride(1),
                    T, B, V, targets.data_ptr(), tstride, off_ptr, input_lengths.data_ptr(), target_lengths.data_ptr(),
                    lmax, int(blank), 0, nll.data_ptr(), None, 0, st), "avctc_ctc_forward")
                _lib.check(L.avctc_ctc_reduce(nll.data_ptr(), target_lengths.data_ptr(), B, red, int(zero_infinity),
                                              out.data_ptr(), st), "avctc_ctc_reduce")
        ctx.grad = grad
        loss = out if red == 0 else out.reshape(())
        return loss.to(log_probs.dtype) if log_probs.dtype != torch.float32 else loss

    @staticmethod
    def backward(ctx, grad_out):
        grad, ctx.grad = ctx.grad, None
        if grad is None:
            # The gradient was computed at forward time and handed to autograd by the first backward WITHOUT keeping a
            # reference (a second owner would make AccumulateGrad clone the [T,B,V] tensor: +50 % HBM traffic).
            raise RuntimeError("CTC loss: backward through this graph a second time is not supported (the gradient "
                               "computed at forward time was consumed by the first backward); call the loss again")
        T, B, V = grad.shape
        dev = grad.device
        go = grad_out.detach().to(torch.float32).contiguous()
        gstride = 0 if go.numel() == 1 else 1
        with _lib.device_guard(dev):
            _lib.check(_lib.lib().avctc_ctc_scale_grad(grad.data_ptr(), _lib.dtype_enum(grad), T, B, V, go.data_ptr(),
                                                       gstride, _lib.stream_ptr(dev)), "avctc_ctc_scale_grad")
        return grad, None, None, None, None, None, None


_UNIT = {}


def _unit(dev):
    """fp32 1.0 on `dev`, written once (long before any kernel that reads it is enqueued)."""
    t = _UNIT.get(dev)
    if t is None:
        t = _UNIT[dev] = torch.ones(1, dtype=torch.float32, device=dev)
    return t


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False):
    """Functional form, same argument order as torch.nn.functional.ctc_loss."""
    if reduction not in _lib.REDUCTION:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    unbatched = log_probs.dim() == 2
    if unbatched:       # (T,V) like torch; the unsqueeze stays OUTSIDE the Function so autograd squeezes the gradient back
        log_probs = log_probs.unsqueeze(1)
        targets = torch.as_tensor(targets)
        targets = targets.unsqueeze(0) if targets.dim() == 1 else targets
        input_lengths = torch.as_tensor(input_lengths).reshape(1)
        target_lengths = torch.as_tensor(target_lengths).reshape(1)
    out = _CTCLossFn.apply(log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity)
    if unbatched and reduction == "none":
        out = out.reshape(())
    return out


class CTCLoss(nn.Module):
    """Same constructor and forward as torch.nn.CTCLoss (reference use: trainer.py:25, decoder.py:12)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _lib.REDUCTION:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity)

    def extra_repr(self):
        return f"blank={self.blank}, reduction={self.reduction}, zero_infinity={self.zero_infinity}"
