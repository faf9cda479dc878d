"""CTC fwd+bwd (config 2): depth of the scan's row ring (ctc_ng = row groups in flight per producer warp; fewer groups =
less shared memory per scan CTA = more gradient CTAs resident beside it while the scan runs).
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
    for ov, idle in ((4, 0), (4, 4), (4, 8), (3, 4), (3, 8), (2, 4), (2, 8), (2, 16)):
        _lib.set_tuning("ctc_ng", ov); _lib.set_tuning("ctc_l2pf", idle); _lib.set_tuning("ctc_stamp", 0)
        t_all, _ = bench.event_time(lambda: (fwd(), bwd()), 20, 3, flush, dev)
        t_b, _ = bench.event_time(bwd, 10, 2, flush, dev)
        t_f, _ = bench.event_time(fwd, 10, 2, flush, dev)
        _lib.set_tuning("ctc_stamp", 1)
        flush(); fwd(); bwd(); torch.cuda.synchronize()
        blk = ws[wsb - ((256 + 4 * B + 255) // 256) * 256:][64:96].cpu().numpy().view(np.uint64)
        t0 = np.uint64(~blk[0]); scan_end = (int(blk[1]) - int(t0)) / 1e3; grad_end = (int(blk[2]) - int(t0)) / 1e3
        first = (int(np.uint64(~blk[3])) - int(t0)) / 1e3 if blk[3] else float("nan")
        g = grad.clone()
        if ref is None:
            ref = g
        dmax = float((g - ref).abs().max() / ref.abs().max())
        print(f"T={T} ctc_ng={ov} l2pf={idle}: fwd+bwd {t_all*1e3:.1f} us, fwd alone {t_f*1e3:.1f} us, bwd alone {t_b*1e3:.1f} us | stamps: scan end {scan_end:.1f} us, "
              f"first early chunk {first:.1f} us, grad end {grad_end:.1f} us | grad vs first config {dmax:.2e}", flush=True)
