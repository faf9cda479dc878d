"""GEMM micro-benchmark: back-to-back launches + phase timestamps of CTA (0,0,0)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_av_model_b200 import _lib
from multimodal_av_model_b200.gemm import gemm, operand
dev = torch.device("cuda:0")
L = _lib.lib()
def run(M, N, K, reps=20, out_dtype=torch.bfloat16):
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    for _ in range(3): gemm(operand(x), operand(w), M, N, K, out)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gemm(operand(x), operand(w), M, N, K, out)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / reps * 1e3
    _lib.set_tuning("gemm_dbg", 1)
    gemm(operand(x), operand(w), M, N, K, out); torch.cuda.synchronize()
    _lib.set_tuning("gemm_dbg", 0)
    ts = (ctypes.c_longlong * 16)()
    L._cdll.avctc_debug_gemm_timestamps(ts)
    t0 = ts[0]
    names = ["start", "setup done", "first full", "last full", "acc ready", "epi done", "exit sync", "staged", "bar"]
    n_ctas = ((M + 127) // 128) * ((N + 127) // 128)
    if n_ctas <= 2048:
        buf = (ctypes.c_longlong * (2 * n_ctas))()
        L._cdll.avctc_debug_gemm_cta_times(buf, n_ctas)
        st = [buf[2 * i] for i in range(n_ctas)]; en = [buf[2 * i + 1] for i in range(n_ctas)]
        t00 = min(st)
        print(f"   CTAs {n_ctas}: start spread {(max(st)-t00)/1e3:.2f} us, kernel span {(max(en)-t00)/1e3:.2f} us, CTA life min/med/max "
              f"{min(e-s for s,e in zip(st,en))/1e3:.2f}/{sorted(e-s for s,e in zip(st,en))[n_ctas//2]/1e3:.2f}/{max(e-s for s,e in zip(st,en))/1e3:.2f} us us")
    print(f"M={M} N={N} K={K}: {t:.1f} us/launch, {2*M*N*K/t/1e6:.0f} TFLOP/s | " + "  ".join(f"{n}+{(ts[i]-t0)/1e3:.2f}" for i, n in enumerate(names)), flush=True)
run(4800, 512, 512); run(4800, 512, 1024); run(4800, 1024, 512); run(512, 512, 4800, out_dtype=torch.float32); run(8192, 8192, 8192, reps=3); run(128, 128, 64); run(128, 128, 4096)
