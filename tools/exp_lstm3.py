import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200.fusion_module import _BiLSTMFn
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, H = 8, 150, 512
ref = torch.nn.LSTM(H, H, num_layers=2, batch_first=True, bidirectional=True).to(dev)
x = torch.randn(B, T, H, device=dev, dtype=torch.bfloat16)
pkg._lib.set_tuning("lstm_dbg", 1)
with torch.no_grad():
    for _ in range(2): _BiLSTMFn.apply(x, *ref._flat_weights)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
pkg._lib.lib()._cdll.avctc_debug_lstm_phases(buf)
names = ["prefetch issue", "cluster wait", "mma+red+sync", "pointwise+sync", "dsmem stores", "arrive", "deferred stores", "-"]
tot = sum(buf[:7])
for n, v in zip(names, buf):
    print(f"{n:18s} {v / T:9.0f} cycles/step")
print("total", tot / T, "cycles/step")
