"""Beam-search decode with the reference's call surface, backed by the batched sm_100a kernel.

    simple_beam_search(log_probs[T,V], beam_width=5, blank=0) -> list[int]   /root/reference/beam_search.py:2
    fast_decode(ids, tokenizer) -> str                                       /root/reference/beam_search.py:45
    beam_search_batch(log_probs[N,T,V], ...) -> list[list[int]]              new: one launch for a whole batch

The reference decodes one utterance at a time from Python (model/trainer.py:229-242); the batched entry
decodes a whole batch with one C call (csrc/beam_search.cu: one fused persistent kernel whose top-k warps feed
recurrence warps through shared memory for up to 16 x SMs short utterances, a top-k pass over all N*T rows + one warp
per utterance above that) and syncs once, only because a Python list is returned.  Token lists are bit-exact with the
reference on CPU, including torch.topk's tie order.
"""
from __future__ import annotations

import torch

from . import _lib


def beam_search_batch(log_probs, beam_width: int = 5, blank: int = 0, lengths=None, return_debug: bool = False):
    """log_probs: CUDA [N,T,V] (any float dtype; read as fp32, last dim contiguous).
    lengths: optional int64 [N] frames to decode per utterance (the reference decodes all T padded
    frames, trainer.py:230, which is the default).  Returns N token lists; with return_debug also
    (final beam scores [N,beam] float64, raw beam paths [N,beam,T] int32)."""
    if log_probs.dim() != 3:
        raise RuntimeError("log_probs must be (N, T, V)")
    if not log_probs.is_cuda:
        return _beam_search_from_host(log_probs, beam_width, blank, lengths, return_debug)
    lp = log_probs.detach()
    if lp.dtype != torch.float32:
        lp = lp.float()
    if lp.stride(2) != 1:
        lp = lp.contiguous()
    N, T, V = lp.shape
    k = int(beam_width)
    if k < 1 or k > V:
        raise RuntimeError("selected index k out of range")   # what torch.topk raises in the reference
    dev = lp.device
    L = _lib.lib()
    ws_bytes = int(L.avctc_beam_workspace_bytes(N, T, V, k))
    if ws_bytes == 0:
        raise RuntimeError(f"beam_width={k} is not supported by the sm_100a beam kernel (max 32)")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_ids = torch.empty((N, max(T, 1)), dtype=torch.int32, device=dev)
    out_len = torch.zeros(N, dtype=torch.int32, device=dev)
    dbg_s = torch.zeros((N, k), dtype=torch.float64, device=dev) if return_debug else None
    dbg_p = torch.zeros((N, k, max(T, 1)), dtype=torch.int32, device=dev) if return_debug else None
    if lengths is not None:
        lengths = torch.as_tensor(lengths).to(device=dev, dtype=torch.long).contiguous()
    with _lib.device_guard(dev):
        _lib.check(L.avctc_beam_search(
            lp.data_ptr(), lp.stride(0), lp.stride(1), N, T, V,
            lengths.data_ptr() if lengths is not None else None, k, int(blank),
            out_ids.data_ptr(), out_len.data_ptr(),
            dbg_s.data_ptr() if return_debug else None, dbg_p.data_ptr() if return_debug else None,
            ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev)), "avctc_beam_search")
    status = ws[:4].view(torch.int32)
    packed = torch.cat([out_len, status]).cpu()    # the one host sync
    if int(packed[-1]) != 0:
        raise RuntimeError("beam search: tied top-k values with beam_width*64 > V need torch.topk's "
                           "nth_element order, which the sm_100a kernel does not reproduce")
    lens = packed[:-1].tolist()
    ids = out_ids.cpu().numpy()
    if N <= 4:
        res = [ids[i, :lens[i]].tolist() for i in range(N)]
    else:       # one flat conversion + list slicing (thousands of per-row tensor slices are ~10 us each)
        import numpy as np
        valid = np.arange(ids.shape[1])[None, :] < np.asarray(lens)[:, None]
        flat = ids[valid].tolist()
        res, o = [], 0
        for l in lens:
            res.append(flat[o:o + l])
            o += l
    if return_debug:
        return res, dbg_s.cpu(), dbg_p.cpu()
    return res


def _beam_search_from_host(log_probs, beam_width, blank, lengths, return_debug, chunk_bytes=32 << 20, device=None):
    """Host-resident log-probs (the reference accepts any device): streamed to the GPU in ~32 MB chunks of whole
    utterances on a copy stream, two device buffers, so that the decode of chunk i runs under the transfer of chunk
    i+1 and the call is bound by the PCIe copy alone.  Nothing is decoded on the host."""
    if not torch.cuda.is_available():
        raise RuntimeError("beam_search_batch: the sm_100a beam kernel needs a CUDA device (no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lp = log_probs.detach()
    if lp.dtype != torch.float32:
        lp = lp.float()
    lp = lp.contiguous()
    N, T, V = lp.shape
    per = max(1, T * V * 4)
    step = max(1, min(N, chunk_bytes // per)) if N else 1
    if N <= step:
        return beam_search_batch(lp.to(dev, non_blocking=True), beam_width, blank,
                                 None if lengths is None else torch.as_tensor(lengths), return_debug)
    if lengths is not None:
        lengths = torch.as_tensor(lengths).to(torch.long)
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(device=dev)
    bufs = [torch.empty((step, T, V), dtype=torch.float32, device=dev) for _ in range(2)]
    free = [None, None]                # event: the decode that last read buffer j has been enqueued and finished
    res, dbg_s, dbg_p = [], [], []
    pending = None                     # (buffer index, n, lengths slice, copy-done event)
    starts = list(range(0, N, step))

    def issue(ci):
        j = ci % 2
        n = min(step, N - starts[ci])
        with torch.cuda.stream(copy):
            if free[j] is not None:
                copy.wait_event(free[j])
            bufs[j][:n].copy_(lp[starts[ci]:starts[ci] + n], non_blocking=True)
            ev = copy.record_event()
        return (j, n, None if lengths is None else lengths[starts[ci]:starts[ci] + n], ev)

    pending = issue(0)
    for ci in range(len(starts)):
        j, n, ln, ev = pending
        pending = issue(ci + 1) if ci + 1 < len(starts) else None          # next transfer first, then this decode
        main.wait_event(ev)
        out = beam_search_batch(bufs[j][:n], beam_width, blank, ln, return_debug)   # syncs on this chunk only
        free[j] = main.record_event()
        if return_debug:
            res.extend(out[0]); dbg_s.append(out[1]); dbg_p.append(out[2])
        else:
            res.extend(out)
    if return_debug:
        return res, torch.cat(dbg_s), torch.cat(dbg_p)
    return res


def simple_beam_search(log_probs: torch.Tensor, beam_width=5, blank=0):
    """Same signature and result as the reference: (T, V) log-probabilities -> collapsed token ids."""
    if log_probs.dim() != 2:
        raise ValueError("not enough values to unpack (expected 2)")   # T, V = log_probs.shape
    return beam_search_batch(log_probs.unsqueeze(0), beam_width=beam_width, blank=blank)[0]


def fast_decode(ids, tokenizer):
    """ids -> text (beam_search.py:45-49): drop blank / out-of-range ids, U+2581 -> space, strip."""
    pieces = tokenizer.id_to_token
    n, blank = tokenizer.vocab_size, tokenizer.blank_id
    return "".join(pieces[i] for i in ids if i != blank and 0 <= i < n).replace("▁", " ").strip()
