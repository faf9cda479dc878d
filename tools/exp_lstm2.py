import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200.fusion_module import _BiLSTMFn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, H = 8, 150, 512
ref = torch.nn.LSTM(H, H, num_layers=2, batch_first=True, bidirectional=True).to(dev)
x = torch.randn(B, T, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
r = torch.randn(B, T, 2 * H, device=dev, dtype=torch.bfloat16)
def ours():
    x.grad = None
    y = _BiLSTMFn.apply(x, *ref._flat_weights); y.backward(r)
for mode in (0,):
    pass
    for _ in range(3): ours()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ours(); torch.cuda.synchronize()
    print("lstm_cluster =", mode)
    for e in prof.key_averages():
        if "lstm" in e.key or "gemm" in e.key:
            print(f"   {e.key[:70]:70s} n={e.count:3d} total {e.device_time_total:9.1f} us")
