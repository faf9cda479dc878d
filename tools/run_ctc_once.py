"""One CTC forward+backward at config 2 (B=64, T=1000, V=801) through the C ABI — the process ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
r = bench.bench_ctc(dev, iters=1, Ts=(1000,))
print(r)
