"""Host-side cost of one hot-path step (fusion -> BiLSTM -> head -> CTC + InfoNCE, forward + backward): cProfile of the
Python/ctypes/autograd enqueue work while the GPU runs behind (the hot path is host-enqueue bound)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_av_model_b200.synthetic import make_features
dev = torch.device("cuda:0")
tr = bench.build_models(dev, encoders=False)
f = make_features(pairs=8, t_v=150, t_enc=249, seed=1234, dtype=torch.bfloat16)
fd = {k: [t.to(dev) for t in v] for k, v in f.items()}
for k in ("audio", "middle"):
    fd[k] = [t.requires_grad_() for t in fd[k]]
def hot_step():
    tr.optimizer.zero_grad(set_to_none=True)
    for k in ("audio", "middle"):
        for t in fd[k]:
            t.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        total = tr.hot_path_loss(fd["visual"], fd["audio"], fd["middle"], fd["masks"], fd["texts"], fd["lens"])[0]
    total.backward()
for _ in range(5): hot_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): hot_step()
t_host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 20 * 1e3
print(f"host enqueue {t_host:.3f} ms/step, with final sync {t_all:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20): hot_step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
