import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200.fusion_module import _BiLSTMFn
dev = torch.device("cuda:0")
torch.manual_seed(0)
for B in (8, 32):
    T, H = 150, 512
    ref = torch.nn.LSTM(H, H, num_layers=2, batch_first=True, bidirectional=True).to(dev)
    x = torch.randn(B, T, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(B, T, 2 * H, device=dev, dtype=torch.bfloat16)
    def ours():
        x.grad = None
        y = _BiLSTMFn.apply(x, *ref._flat_weights); y.backward(r)
    def ours_fwd():
        with torch.no_grad(): _BiLSTMFn.apply(x, *ref._flat_weights)
    def cudnn():
        x.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y, _ = ref(x)
        y.backward(r)
    def cudnn_fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16): ref(x)
    for name, fn in (("ours fwd", ours_fwd), ("ours fwd+bwd", ours), ("cudnn fwd", cudnn_fwd), ("cudnn fwd+bwd", cudnn)):
        for _ in range(3): fn()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        print(f"B={B} {name}: {a.elapsed_time(b)/10:.3f} ms (wall {(time.perf_counter()-t0)*100:.3f} ms)", flush=True)
