"""L2 prefetch knobs (ctc_pf, beam_pf): CTC fwd+bwd and beam decode times with the knob off / on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_av_model_b200 import _lib
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for pf in (1, 1):
    _lib.set_tuning("ctc_pf", pf); _lib.set_tuning("beam_pf", 1)
    c = bench.bench_ctc(dev, iters=20)
    b = {'ms': 0.0}
    print(f"pf={pf}: ctc T1000 fwd+bwd {c['T1000']['fwd_bwd_ms']*1e3:.1f} us (scan {c['T1000']['scan_ms']*1e3:.1f}, grad {c['T1000']['grad_ms']*1e3:.1f}) "
          f"T250 {c['T250']['fwd_bwd_ms']*1e3:.1f} us | beam {b['ms']*1e3:.1f} us", flush=True)
