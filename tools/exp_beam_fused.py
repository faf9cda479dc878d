"""A/B of the beam decode kernels on one B200: two-phase (top-k pass, then recurrence) against the fused kernel and
the default policy, at config 5 (4096 x [150, 800], beam 10) and at smaller batches (where the auto policy must switch).
Each variant's token lists are compared with the two-phase result before it is timed.

    python tools/exp_beam_fused.py [out.txt]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import multimodal_av_model_b200 as pkg  # noqa: E402
from multimodal_av_model_b200 import _lib  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    dev = torch.device("cuda:0")
    T, V, beam, blank = 150, 800, 10, 3
    L = _lib.lib()
    flush = bench.l2_flusher(dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)

    peak = bench.measured_peaks()["hbm"]
    g = torch.Generator(device="cuda").manual_seed(7)
    lp_all = (3 * torch.randn(4096, T, V, generator=g, device=dev)).log_softmax(-1)
    for N in (4096, 3072, 2048, 1024, 512, 64, 16):
        lp = lp_all[:N]
        wsb = int(L.avctc_beam_workspace_bytes(N, T, V, beam))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        out = torch.empty((N, T), dtype=torch.int32, device=dev)
        ol = torch.empty(N, dtype=torch.int32, device=dev)

        def run():
            _lib.check(L.avctc_beam_search(lp.data_ptr(), lp.stride(0), lp.stride(1), N, T, V, None, beam, blank,
                                           out.data_ptr(), ol.data_ptr(), None, None, ws.data_ptr(), wsb, st), "beam")
        variants = [("two-phase", 0), ("fused", 1), ("auto", -1)]
        want = None
        for name, fused in variants:
            _lib.set_tuning("beam_fused", fused)
            out.zero_(); ol.zero_()
            run()
            torch.cuda.synchronize(dev)
            got = (out.clone(), ol.clone())
            if want is None:
                want = got
            same = bool(torch.equal(got[1], want[1])) and all(
                torch.equal(got[0][i, :int(want[1][i])], want[0][i, :int(want[1][i])]) for i in range(0, N, max(1, N // 257)))
            mean, best = bench.event_time(run, 10 if N >= 1024 else 20, 3, flush, dev)
            gb = N * T * V * 4 / 1e9
            say(f"N={N:5d}  {name:12s}  mean {mean * 1e3:8.1f} us  min {best * 1e3:8.1f} us  {gb / mean * 1e3:7.1f} GB/s  "
                f"frac {gb / mean * 1e3 / peak:.3f}  identical={same}")
        _lib.set_tuning("beam_fused", -1)
    if out_path:
        with open(out_path, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
