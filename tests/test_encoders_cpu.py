"""CPU: host-side rewrites inside the producer encoders change no value.

* VisualEncoder._frontend_as_2d == frontend3D (same parameters; Conv3d over one input channel as a 7x7 Conv2d over the
  five temporal taps, BatchNorm3d as BatchNorm2d over B*T frames, MaxPool3d((1,3,3)) as MaxPool2d) incl. running stats
* the audio feature-extractor cache returns the same tensor for an unchanged input and recomputes for a changed one
* the frozen feature extractor no longer asks autograd for a gradient (HF's gradient-checkpointing aid)
"""
import torch

import multimodal_av_model_b200 as pkg
from multimodal_av_model_b200.encoders import unfreeze_middle_layers, xlsr_large_config


def test_visual_frontend_2d_equals_3d_train_and_eval():
    torch.manual_seed(0)
    a, b = pkg.VisualEncoder(), pkg.VisualEncoder()
    b.load_state_dict(a.state_dict())
    x = torch.rand(2, 1, 9, 96, 96)
    for mode in ("train", "eval"):
        getattr(a, mode)(); getattr(b, mode)()
        y3 = a.frontend3D(x)
        bb, c, t, h, w = y3.shape
        y2 = b._frontend_as_2d(x)
        assert y2.shape == (bb * t, c, h, w)
        assert torch.allclose(y3.transpose(1, 2).reshape(bb * t, c, h, w), y2, atol=2e-5, rtol=1e-5)
    bn_a, bn_b = a.frontend3D[1], b.frontend3D[1]
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, atol=1e-6)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, atol=1e-6)
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 1


def _tiny_audio():
    torch.manual_seed(0)
    cfg = xlsr_large_config(num_hidden_layers=10, hidden_size=64, num_attention_heads=4, intermediate_size=128,
                            num_conv_pos_embedding_groups=4)
    return pkg.AudioEncoder(freeze=True, config=cfg)


def test_audio_feature_cache_and_no_grad_through_frozen_extractor():
    aud = _tiny_audio()
    unfreeze_middle_layers(aud.model)
    aud.eval()
    x = 0.1 * torch.randn(2, 8000)
    m = torch.ones(2, 8000, dtype=torch.bool)
    calls = []
    inner = aud.model.feature_extractor.conv_layers[0].register_forward_hook(lambda *_: calls.append(1))
    a1, mid1 = aud(x, attention_mask=m)
    a2, _ = aud(x, attention_mask=m)
    assert torch.equal(a1, a2) and len(calls) == 1                 # second call reused the conv features
    a3, _ = aud(x.clone(), attention_mask=m)
    assert torch.allclose(a1, a3) and len(calls) == 2              # different tensor object: recomputed
    x.add_(0.01)
    a4, _ = aud(x, attention_mask=m)
    assert len(calls) == 3 and not torch.allclose(a1, a4)          # in-place change bumps the version: recomputed
    inner.remove()
    assert aud.model.feature_extractor._requires_grad is False
    aud.train()
    out, mid = aud(x, attention_mask=m)
    (out.sum() + mid.sum()).backward()
    grads = {n for n, p in aud.named_parameters() if p.grad is not None}
    assert grads and all("encoder.layers." in n for n in grads)


def test_audio_feature_cache_is_keyed_on_object_identity_not_address():
    """A new batch is often allocated at the address of the freed previous one with the same version counter (0):
    the cache must still recompute."""
    aud = _tiny_audio()
    aud.eval()
    m = torch.ones(1, 4000, dtype=torch.bool)
    outs = []
    for seed in (1, 2, 3):
        x = torch.randn(1, 4000, generator=torch.Generator().manual_seed(seed)) * 0.1      # same shape, often same address
        outs.append(aud(x, attention_mask=m)[0].clone())
        del x
    assert not torch.allclose(outs[0], outs[1]) and not torch.allclose(outs[1], outs[2])
