"""Per-kernel device timing through the C ABI with preallocated buffers (no Python/allocator overhead)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200 import _lib

dev = torch.device("cuda:0")
L = _lib.lib()
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, iters=10, warm=3, do_flush=True):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        if do_flush: flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def ctc(T, B=64, V=801, blank=0, lmin=10, lmax=80, dtype=torch.float32, k=0):
    g = torch.Generator().manual_seed(T)
    lp = torch.randn(T, B, V, generator=g).log_softmax(-1).to(dtype).to(dev)
    rng = np.random.default_rng(T)
    hi = min(lmax, T // 2 - 1)
    tl = rng.integers(min(lmin, hi), hi + 1, size=B); Lm = int(tl.max())
    il = rng.integers(max(2 * Lm + 1, T // 2), T + 1, size=B)
    tg = torch.from_numpy(rng.integers(1, V, size=(B, Lm))).to(dev)
    il = torch.from_numpy(il).to(dev); tl = torch.from_numpy(tl).to(dev)
    _lib.set_tuning("ctc_k", k)
    wsb = L.avctc_ctc_workspace_bytes(T, B, Lm)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    nll = torch.empty(B, device=dev); go = torch.ones(1, device=dev)
    grad = torch.empty_like(lp)
    dt = _lib.dtype_enum(lp)
    def fwd(ng):
        _lib.check(L.avctc_ctc_forward(lp.data_ptr(), dt, lp.stride(0), lp.stride(1), T, B, V, tg.data_ptr(), tg.stride(0), None,
                            il.data_ptr(), tl.data_ptr(), Lm, blank, ng, nll.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
    def bwd():
        _lib.check(L.avctc_ctc_backward(lp.data_ptr(), dt, lp.stride(0), lp.stride(1), T, B, V, tg.data_ptr(), tg.stride(0), None,
                             il.data_ptr(), tl.data_ptr(), Lm, blank, 1, 1, nll.data_ptr(), go.data_ptr(), 0, grad.data_ptr(),
                             ws.data_ptr(), wsb, st), "bwd")
    t_a = timeit(lambda: fwd(0)); t_ab = timeit(lambda: fwd(1)); t_g = timeit(bwd)
    t_all = timeit(lambda: (fwd(1), bwd()))
    alg = 2 * T * B * V * lp.element_size()
    _lib.set_tuning("ctc_k", 0)
    print(f"ctc T={T} B={B} V={V} Lmax={Lm} {str(dtype)[6:]} k={k}: alpha {t_a*1e3:.0f}us  alpha||beta {t_ab*1e3:.0f}us ({t_ab*1e6/T:.0f} ns/frame)  "
          f"grad {t_g*1e3:.0f}us ({alg/t_g/1e6:.0f} GB/s)  fwd+bwd {t_all*1e3:.0f}us = {alg/t_all/1e6:.0f} GB/s  ws {wsb/1e6:.0f} MB", flush=True)

def beam(N=4096, T=150, V=800, k=10, fast=1, scale=3.0):
    g = torch.Generator().manual_seed(7)
    lp = (scale * torch.randn(N, T, V, generator=g)).log_softmax(-1).to(dev)
    _lib.set_tuning("beam_fast", fast)
    wsb = L.avctc_beam_workspace_bytes(N, T, V, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty((N, T), dtype=torch.int32, device=dev); ol = torch.empty(N, dtype=torch.int32, device=dev)
    def run():
        _lib.check(L.avctc_beam_search(lp.data_ptr(), lp.stride(0), lp.stride(1), N, T, V, None, k, 3, out.data_ptr(), ol.data_ptr(),
                            None, None, ws.data_ptr(), wsb, st), "beam")
    t = timeit(run, iters=5)
    _lib.set_tuning("beam_fast", 1)
    print(f"beam N={N} T={T} V={V} k={k} fast={fast}: {t:.3f} ms  {N/t*1e3:.0f} utt/s  {N*T*V*4/t/1e6:.0f} GB/s", flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "ctc"):
        for T in (75, 250, 1000):
            ctc(T)
        ctc(1000, k=4)
        ctc(1000, lmin=10, lmax=30)
        ctc(1000, lmin=40, lmax=60)
        ctc(1000, dtype=torch.bfloat16)
        ctc(150, B=8, V=800, blank=3, lmin=20, lmax=58)
    if what == "ctc1000":
        ctc(1000)
    if what == "beam1":
        beam(k=10)
    if what in ("all", "beam"):
        beam(k=10); beam(k=10, fast=0); beam(k=5); beam(k=5, fast=0); beam(N=64, k=10)
