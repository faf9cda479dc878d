"""Quick device-side timing probe (not the bench contract): CTC config 2 and beam config 5."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_av_model_b200 as pkg

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

def ctc_case(T, B=64, V=801, blank=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(T)
    lp = torch.randn(T, B, V, generator=g).log_softmax(-1).to(dtype).to(dev)
    rng = np.random.default_rng(T)
    hi = min(80, T // 2 - 1)
    tl = rng.integers(10, hi + 1, size=B); Lm = int(tl.max())
    il = rng.integers(max(2 * Lm + 1, T // 2), T + 1, size=B)
    tg = rng.integers(1, V, size=(B, Lm))
    return lp, torch.from_numpy(tg).to(dev), torch.from_numpy(il).to(dev), torch.from_numpy(tl).to(dev)

res = {}
for T in (75, 150, 250, 500, 1000):
    lp, tg, il, tl = ctc_case(T)
    x = lp.clone().requires_grad_()
    def ours():
        x.grad = None
        pkg.ctc_loss(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
    def ours_fwd():
        with torch.no_grad(): pkg.ctc_loss(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
    def theirs():
        x.grad = None
        torch.nn.functional.ctc_loss(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
    m, mn = timeit(ours); mf, _ = timeit(ours_fwd); tm, tmn = timeit(theirs)
    alg = 2 * T * 64 * 801 * 4
    res[f"ctc_T{T}"] = dict(ours_ms=m, ours_min_ms=mn, fwd_only_ms=mf, torch_ms=tm, GBs=alg / m / 1e6, torch_GBs=alg / tm / 1e6)
    print(T, res[f"ctc_T{T}"], flush=True)
    for k in (2, 4):
        pkg._lib.set_tuning("ctc_k", k)
        mk, _ = timeit(ours)
        print("   ctc_k", k, mk, flush=True)
    pkg._lib.set_tuning("ctc_k", 0)
    if T == 1000:
        xb = lp.bfloat16().requires_grad_()
        def ours_bf():
            xb.grad = None
            pkg.ctc_loss(xb, tg, il, tl, blank=0, reduction="mean", zero_infinity=True).backward()
        mb, _ = timeit(ours_bf)
        print("   bf16", mb, 2 * T * 64 * 801 * 2 / mb / 1e6, "GB/s", flush=True)

g = torch.Generator().manual_seed(7)
N, T, V = 4096, 150, 800
lp = (3 * torch.randn(N, T, V, generator=g)).log_softmax(-1).to(dev)
for beam in (10, 5):
    for fast in (1, 0):
        pkg._lib.set_tuning("beam_fast", fast)
        m, mn = timeit(lambda: pkg.beam_search_batch(lp, beam_width=beam, blank=3), iters=5, warm=2)
        print("beam", beam, "fast", fast, m, "ms", N / m * 1e3, "utt/s (incl D2H + list build)", flush=True)
pkg._lib.set_tuning("beam_fast", 1)
json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)
