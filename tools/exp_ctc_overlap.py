"""CTC fwd+bwd (config 2): gradient pass starting per finished sample (ctc_overlap) vs waiting for the scan grid.
Prints CUDA-event times and, from the kernels' own globaltimer stamps (ctc_stamp), when the scan ended and when the
gradient pass ended relative to the first scan CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from multimodal_av_model_b200 import _lib
dev = torch.device("cuda:0")
L = _lib.lib()
flush = bench.l2_flusher(dev)
st = torch.cuda.current_stream(dev).cuda_stream
for T in (250, 1000):
    lp, tg, il, tl, Lm = bench.ctc_case(T, dev)
    B, V = lp.shape[1], lp.shape[2]
    wsb = L.avctc_ctc_workspace_bytes(T, B, Lm)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, device=dev); go = torch.ones(1, device=dev); grad = torch.empty_like(lp); loss = torch.empty(1, device=dev)
    def fwd():
        _lib.check(L.avctc_ctc_forward(lp.data_ptr(), 0, lp.stride(0), lp.stride(1), T, B, V, tg.data_ptr(), tg.stride(0),
                                       None, il.data_ptr(), tl.data_ptr(), Lm, 0, 1, nll.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
        _lib.check(L.avctc_ctc_reduce(nll.data_ptr(), tl.data_ptr(), B, 1, 1, loss.data_ptr(), st), "reduce")
    def bwd():
        _lib.check(L.avctc_ctc_backward(lp.data_ptr(), 0, lp.stride(0), lp.stride(1), T, B, V, tg.data_ptr(), tg.stride(0),
                                        None, il.data_ptr(), tl.data_ptr(), Lm, 0, 1, 1, nll.data_ptr(), go.data_ptr(), 0,
                                        grad.data_ptr(), ws.data_ptr(), wsb, st), "bwd")
    ref = None
    for ov in (0, 1, 0, 1):
        _lib.set_tuning("ctc_overlap", ov); _lib.set_tuning("ctc_stamp", 0)
        t_all, _ = bench.event_time(lambda: (fwd(), bwd()), 20, 3, flush, dev)
        t_b, _ = bench.event_time(bwd, 10, 2, flush, dev)
        _lib.set_tuning("ctc_stamp", 1)
        flush(); fwd(); bwd(); torch.cuda.synchronize()
        blkall = ws[wsb - ((256 + 4 * B + 255) // 256) * 256:][64:192].cpu().numpy().view(np.uint64)
        blk = blkall[:4]; dur = blkall[4:8]
        t0 = np.uint64(~blk[0]); scan_end = (int(blk[1]) - int(t0)) / 1e3; grad_end = (int(blk[2]) - int(t0)) / 1e3
        first = (int(np.uint64(~blk[3])) - int(t0)) / 1e3 if blk[3] else float("nan")
        hist = ws[wsb - ((256 + 4 * B + 255) // 256) * 256:][176:240].cpu().numpy().view(np.int32)
        tick = ws[wsb - ((256 + 4 * B + 255) // 256) * 256:][240:248].cpu().numpy().view(np.int32)
        real_rows = int(il.sum().item()); chunk = max(1, min(16, real_rows // (444 * 8 * 4)))
        g = grad.clone()
        if ref is None:
            ref = g
        dmax = float((g - ref).abs().max() / ref.abs().max())
        print(f"T={T} overlap={ov}: fwd+bwd {t_all*1e3:.1f} us, bwd alone {t_b*1e3:.1f} us | stamps: scan end {scan_end:.1f} us, "
              f"first early chunk {first:.1f} us, grad end {grad_end:.1f} us | recurrence warp: prologue {int(dur[0])/1e3:.1f} us, epilogue {int(dur[1])/1e3:.1f} us, "
              f"slowest {int(dur[2])/1e3:.1f} ns/frame, longest loop {int(dur[3])/1e3:.1f} us | gradient warps: set-up + zero rows {int(blkall[10])/1e3:.1f} us, "
              f"last warp at its rows {(int(blkall[11]) - int(t0))/1e3:.1f} us, last CTA entry {(int(blkall[12]) - int(t0))/1e3:.1f} us | grad vs first config {dmax:.2e}", flush=True)
        print("      gradient CTA entries per 16 us:", " ".join(str(int(v)) for v in hist),
              f"| tickets drawn when the last scan CTA ended: {int(tick[0])} x {chunk} rows of {real_rows} real rows; zero-row tickets {int(tick[1])}", flush=True)
    # the product order: scan, gradient pass, guarded twins, reduction in ONE call
    def fb():
        _lib.check(L.avctc_ctc_forward_backward(lp.data_ptr(), 0, lp.stride(0), lp.stride(1), T, B, V, tg.data_ptr(), tg.stride(0), None,
                                                il.data_ptr(), tl.data_ptr(), Lm, 0, 1, 1, nll.data_ptr(), loss.data_ptr(), go.data_ptr(), 0,
                                                grad.data_ptr(), ws.data_ptr(), wsb, st), "fwd_bwd")
    _lib.set_tuning("ctc_overlap", 1); _lib.set_tuning("ctc_stamp", 0)
    t_fb, _ = bench.event_time(fb, 20, 3, flush, dev)
    fb(); torch.cuda.synchronize()
    print(f"T={T} avctc_ctc_forward_backward: {t_fb*1e3:.1f} us | grad vs first config {float((grad - ref).abs().max() / ref.abs().max()):.2e}", flush=True)
