"""train_epoch variants (prefetch on/off) vs the blocking step loop: per-step host time of train_step / stage,
cudaMalloc count, wall per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_av_model_b200.synthetic import make_batch
dev = torch.device("cuda:0")
tr = bench.build_models(dev)
host = make_batch(pairs=8, seconds=5.0, t_v=150, vocab=bench.VOCAB, seed=1234, pin=True)
for _ in range(3):
    float(tr.train_step(host))
torch.cuda.synchronize()
log = []
_ts, _st = tr.train_step, tr.stage
def ts(b):
    a = time.perf_counter(); r = _ts(b); log.append(("step", (time.perf_counter() - a) * 1e3)); return r
def st(b):
    a = time.perf_counter(); r = _st(b); log.append(("stage", (time.perf_counter() - a) * 1e3)); return r
tr.train_step, tr.stage = ts, st
def mallocs():
    return torch.cuda.memory_stats(dev)["num_device_alloc"]
def blocking(n):
    for _ in range(n):
        float(tr.train_step(host))
def epoch(n):
    tr.train_epoch([host] * n)
    assert tr.last_epoch_steps == n
for name, fn, pf in (("blocking", blocking, True), ("epoch prefetch=0", epoch, False), ("epoch prefetch=1", epoch, True),
                     ("epoch prefetch=1 again", epoch, True), ("epoch prefetch=0 again", epoch, False)):
    tr.prefetch_batches = pf
    torch.cuda.synchronize(); log.clear(); m0 = mallocs()
    t0 = time.perf_counter(); fn(10); torch.cuda.synchronize(); w = (time.perf_counter() - t0) / 10 * 1e3
    print(f"{name}: {w:.1f} ms/step, cudaMallocs {mallocs() - m0}, reserved {torch.cuda.memory_reserved(dev) >> 20} MiB", flush=True)
    print("   step ms:", " ".join(f"{v:.0f}" for k, v in log if k == "step"))
    print("   stage ms:", " ".join(f"{v:.1f}" for k, v in log if k == "stage"), flush=True)
