"""GPU: parity at the sizes bench.py quotes (BASELINE configs 2, 3, 4) and regressions for the round-1 review.

What is compared with what:
  * config 3 (B=32, T_v=150, T_a=249): projections + cross attention fwd/bwd vs oracle/torch_port.FusionPort (fp32 CPU;
    the port is pinned to the reference's own modules at small dims by tests/test_oracle_golden.py — the reference
    modules themselves cannot travel to the GPU box)
  * BiLSTM at (32,150,512) vs torch.nn.LSTM fp32
  * config 4 (8 pairs, T_v=150, T_enc=249, V=800): hot-path loss + parameter gradients vs oracle/torch_port.hot_path_losses
  * config 2 (B=64, T=1000, V=801) with bf16 log-probs vs the float64 C oracle
Tolerances are the MEASURED errors of the kernels with ~2x head-room (bf16 operands, fp32 accumulation), written next
to each assert; north_star's bound is rel 1e-2 on O(1) activations."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


def relerr(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def rms_relerr(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-12)).item()


def test_fusion_projection_config3_full_size_vs_torch_port(record_property):
    from test_fusion_gpu import config3_inputs
    pkg = _pkg()
    torch.manual_seed(0)
    ref = tp.FusionPort(512, 1024, 512)
    ours = pkg.CrossAttentionFusion(512, 1024, 512)
    ours.load_state_dict(ref.state_dict())
    ours.cuda()
    vis, aud, mask = config3_inputs(B=32)
    v1, a1 = vis.clone().requires_grad_(), aud.clone().requires_grad_()
    f_ref, m_ref = ref.projection(v1, a1, mask)
    r = torch.randn_like(f_ref)
    (f_ref * r).sum().backward()
    v2, a2 = vis.cuda().requires_grad_(), aud.cuda().requires_grad_()
    f, m, il = ours.fused_projection(v2, a2, mask.cuda())
    (f * r.cuda()).sum().backward()
    assert torch.equal(m.cpu(), m_ref)
    assert il.cpu().tolist() == (m_ref != 0).sum(1).tolist()
    errs = {"fused": relerr(f, f_ref), "d_audio": relerr(a2.grad, a1.grad), "d_visual": relerr(v2.grad, v1.grad)}
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        if q.grad is None:
            assert p.grad is None
        elif not k.startswith("temporal_model"):
            errs[k] = relerr(p.grad, q.grad)
    print("config-3 bf16 errors (max-norm relative):", {k: round(v, 5) for k, v in errs.items()})
    assert errs["fused"] < 1e-2, errs                 # measured on B200: 3.9e-3
    assert rms_relerr(f, f_ref) < 5e-3
    for k, v in errs.items():
        assert v < 1.5e-2, (k, v, errs)               # measured: 1.6e-3 ... 7.7e-3 (d_audio) over all gradients


def test_bilstm_config3_full_size_vs_torch():
    from multimodal_av_model_b200.fusion_module import _BiLSTMFn
    B, T, H = 32, 150, 512
    torch.manual_seed(7)
    ref = torch.nn.LSTM(H, H, num_layers=2, batch_first=True, bidirectional=True).cuda()
    x = torch.randn(B, T, H, device="cuda")
    r = torch.randn(B, T, 2 * H, device="cuda")
    x1 = x.clone().requires_grad_()
    y_ref, _ = ref(x1)
    (y_ref * r).sum().backward()
    g_ref = [p.grad.clone() for p in ref._flat_weights]
    for p in ref.parameters():
        p.grad = None
    x2 = x.clone().requires_grad_()
    y = _BiLSTMFn.apply(x2, *ref._flat_weights)
    (y.float() * r).sum().backward()
    errs = {"y": relerr(y, y_ref), "dx": relerr(x2.grad, x1.grad)}
    for name, p, g in zip(ref._flat_weights_names, ref._flat_weights, g_ref):
        errs[name] = relerr(p.grad, g)
    print("BiLSTM (32,150,512) bf16 errors:", {k: round(v, 5) for k, v in errs.items()})
    assert errs["y"] < 1e-2, errs                     # measured on B200: 4.7e-3
    for k, v in errs.items():
        assert v < 1e-2, (k, v)                       # measured: 1.8e-3 ... 5.1e-3


class _Enc(torch.nn.Module):
    def forward(self, *a, **k):
        raise AssertionError("encoders are not used by hot_path_loss")


def test_hot_path_step_config4_size_vs_torch_port():
    """8 pairs, T_v=150, T_enc=249, V=800, fp32 features (the fp32 InfoNCE path) and bf16 GEMM operands: loss parts and
    every parameter gradient against the reference's arithmetic on the CPU (trainer.py:98-119)."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_features
    torch.manual_seed(0)
    ref_f, ref_d = tp.FusionPort(512, 1024, 512), tp.DecoderPort(1024, 800, 3)
    proj = torch.nn.Linear(1024, 128)
    fus, dec = pkg.CrossAttentionFusion(512, 1024, 512), pkg.CTCDecoder(1024, 800, blank_id=3)
    fus.load_state_dict(ref_f.state_dict()); dec.load_state_dict(ref_d.state_dict())
    tr = pkg.MultimodalTrainer(_Enc(), _Enc(), fus, dec, CharTokenizer(800), device="cuda")
    tr.projection_layer = torch.nn.Linear(1024, 128).cuda()
    tr.projection_layer.load_state_dict(proj.state_dict())
    f = make_features(pairs=8, t_v=150, t_enc=249, seed=1234)
    feats = [dict(visual=f["visual"][s], audio=f["audio"][s].clone().requires_grad_(), middle=f["middle"][s],
                  mask=f["masks"][s], text=f["texts"][s], text_len=f["lens"][s]) for s in range(2)]
    loss_ref = tp.hot_path_losses(ref_f, ref_d, proj, feats, blank=3)
    loss_ref.backward()
    fd = {k: [t.cuda() for t in v] for k, v in f.items()}
    fd["audio"] = [t.requires_grad_() for t in fd["audio"]]
    total = tr.hot_path_loss(fd["visual"], fd["audio"], fd["middle"], fd["masks"], fd["texts"], fd["lens"])[0]
    total.backward()
    assert abs(total.item() - loss_ref.item()) < 5e-3 * abs(loss_ref.item()), (total.item(), loss_ref.item())
    errs = {}
    for (k, p), (_, q) in list(zip(fus.named_parameters(), ref_f.named_parameters())) + \
            list(zip(dec.named_parameters(), ref_d.named_parameters())):
        if q.grad is None:
            assert p.grad is None, k
        else:
            errs[k] = relerr(p.grad, q.grad)
    for s in range(2):
        errs[f"d_audio{s}"] = relerr(fd["audio"][s].grad, feats[s]["audio"].grad)
    print("config-4 hot-path gradient errors:", {k: round(v, 4) for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 2e-2, (k, v)                       # measured on B200: <= 5.2e-3 (parameters), 8.6e-3 (d_audio)


def test_ctc_bf16_config2_full_size():
    from test_ctc_gpu import make_case, rel
    pkg = _pkg()
    lp, tg, il, tl = make_case(1000, 64, 801, 0, 10, 80, seed=1000, dtype=torch.bfloat16)
    ref = oracle.ctc_loss(lp.float().numpy(), tg, il, tl, blank=0, reduction="mean", zero_infinity=True)
    x = lp.cuda().requires_grad_()
    loss = pkg.ctc_loss(x, torch.from_numpy(tg).cuda(), torch.from_numpy(il).cuda(), torch.from_numpy(tl).cuda(),
                        blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    assert x.grad.dtype == torch.bfloat16
    assert abs(loss.float().item() - ref["loss"]) <= 1e-2 * abs(ref["loss"])
    assert rel(x.grad.float().cpu().numpy(), ref["grad"]) < 1e-2


# ------------------------------------------------------------------------------------------ round-1 review regressions
@pytest.mark.parametrize("B,Tv,Ta,speech", [(2, 64, 30, 9), (3, 40, 12, 1), (2, 150, 20, 20), (1, 7, 5, 1)])
def test_resample_backward_strong_upsampling(B, Tv, Ta, speech):
    """d_audio when one input frame feeds MANY output frames (T_v >= 5 * T', and T' == 1 where every output frame reads
    frame 0): the gather must not cap the number of contributing frames (round-1 advisor finding)."""
    pkg = _pkg()
    torch.manual_seed(B * 10 + Tv)
    aud = torch.randn(B, Ta, 16)
    mask = torch.zeros(B, Ta, dtype=torch.long)
    mask[:, 1:1 + speech] = 1
    if B > 1:
        mask[1, 1:1 + max(1, speech // 2)] = 2
        mask[1, 1 + max(1, speech // 2):] = 3
    a1 = aud.clone().requires_grad_()
    a_ref, m_ref = tp.select_pad_resample(a1, mask, Tv)
    r = torch.randn_like(a_ref)
    (a_ref * r).sum().backward()
    ours = pkg.CrossAttentionFusion(8, 16, 8, num_heads=1).cuda()
    # isolate the resample: identity-free check through the C ABI of the backward kernel
    L = pkg._lib.lib()
    dev = torch.device("cuda")
    st = pkg._lib.stream_ptr(dev)
    xa = torch.empty((B * Tv, 16), dtype=torch.bfloat16, device=dev)
    mo = torch.empty((B, Tv), dtype=torch.long, device=dev)
    il = torch.empty(B, dtype=torch.long, device=dev)
    wsb = int(L.avctc_resample_workspace_bytes(B, Ta))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ac, mc = aud.cuda().contiguous(), mask.cuda().contiguous()
    pkg._lib.check(L.avctc_resample_forward(ac.data_ptr(), 0, mc.data_ptr(), B, Ta, 16, Tv, xa.data_ptr(), mo.data_ptr(),
                                            il.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
    assert torch.equal(mo.cpu(), m_ref)
    assert relerr(xa.view(B, Tv, 16), a_ref) < 1e-2          # bf16 output
    dout = r.to(torch.bfloat16).cuda().contiguous()
    da = torch.empty((B, Ta, 16), dtype=torch.float32, device=dev)
    pkg._lib.check(L.avctc_resample_backward(dout.data_ptr(), B, Ta, 16, Tv, ws.data_ptr(), da.data_ptr(), 0, st), "bwd")
    ref_g = torch.autograd.grad((tp.select_pad_resample(a1, mask, Tv)[0] * dout.float().cpu()).sum(), a1)[0]
    assert (da.cpu() - ref_g).abs().max().item() <= 1e-4 * (ref_g.abs().max().item() + 1e-12)
    del ours


def test_ctc_unbatched_backward():
    """(T,V) log-probs like torch: forward AND backward (the gradient has the input's shape)."""
    pkg = _pkg()
    torch.manual_seed(3)
    lp = torch.randn(40, 20).log_softmax(-1)
    tg = torch.tensor([5, 6, 6, 7])
    x = lp.cuda().requires_grad_()
    y = lp.cuda().requires_grad_()
    a = pkg.ctc_loss(x, tg.cuda(), torch.tensor(40), torch.tensor(4), blank=3, reduction="sum")
    b = torch.nn.functional.ctc_loss(y, tg.cuda(), torch.tensor([40]), torch.tensor([4]), blank=3, reduction="sum")
    a.backward(); b.backward()
    assert x.grad.shape == (40, 20)
    assert torch.allclose(a, b, rtol=1e-4)
    assert (x.grad - y.grad).abs().max().item() < 1e-4 * y.grad.abs().max().item()


def test_evaluate_with_different_lip_lengths():
    """collate_fn pads lip1 and lip2 separately, so T_v1 != T_v2 on real batches (round-1 advisor finding, high)."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_batch
    from test_trainer_gpu import tiny_models
    vis, aud, fus, dec = tiny_models(pkg)
    tr = pkg.MultimodalTrainer(vis, aud, fus, dec, CharTokenizer(800), device="cuda")
    tr.verbose = False
    batches = []
    for s in range(2):
        b = make_batch(pairs=2, seconds=1.0, t_v=30, seed=s, l_range=(3, 8))
        b["lip2"] = b["lip2"][:, :23].contiguous()
        b["lip2_lengths"] = torch.full((2,), 23)
        batches.append(b)
    loss, wer = tr.evaluate(batches)
    assert np.isfinite(loss) and 0.0 <= wer
    l0 = tr.train_epoch(batches)                      # the train step already handled differing shapes
    assert np.isfinite(l0) and tr.last_epoch_steps == 2
