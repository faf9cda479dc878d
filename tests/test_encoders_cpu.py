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


def _run_audio(aud, x, m, sync_free, host_lengths=None, seed=7):
    import numpy as np
    aud.sync_free = sync_free
    torch.manual_seed(seed); np.random.seed(seed)
    return aud(x, attention_mask=m, host_lengths=host_lengths)


def test_audio_sync_free_forward_equals_upstream_forward_eval_and_train():
    """AudioEncoder._forward_sync_free (no GPU read-backs) runs the same submodules with the same random draws as
    Wav2Vec2Model.forward: identical outputs in eval mode and, with the same seeds, in train mode (SpecAugment,
    LayerDrop and dropout all active), for ragged utterance lengths, with and without host-side lengths."""
    for stable in (True, False):
        torch.manual_seed(0)
        cfg = xlsr_large_config(num_hidden_layers=10, hidden_size=64, num_attention_heads=4, intermediate_size=128,
                                num_conv_pos_embedding_groups=4, do_stable_layer_norm=stable,
                                mask_time_prob=0.3, mask_feature_prob=0.1 if stable else 0.0, mask_feature_length=4)
        aud = pkg.AudioEncoder(freeze=True, config=cfg)
        assert aud._sync_free_supported()
        x = 0.1 * torch.randn(3, 12000)
        m = torch.ones(3, 12000, dtype=torch.bool)
        m[1, 9000:] = False; m[2, 5000:] = False
        lens = m.sum(-1)
        for mode in ("eval", "train"):
            getattr(aud, mode)()
            ref = _run_audio(aud, x, m, False)
            for hl in (None, lens):
                got = _run_audio(aud, x, m, True, host_lengths=hl)
                assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1]), (stable, mode)
        aud.train()                                             # nothing padded: upstream drops the attention mask
        full = torch.ones(3, 12000, dtype=torch.bool)
        ref = _run_audio(aud, x, full, False)
        got = _run_audio(aud, x, full, True, host_lengths=full.sum(-1))
        assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
        aud.train()                                             # no attention mask at all
        ref = _run_audio(aud, x, None, False)
        got = _run_audio(aud, x, None, True)
        assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])


def test_audio_prefetch_features_fills_the_cache():
    aud = _tiny_audio()
    aud.train()
    x = 0.1 * torch.randn(2, 8000)
    calls = []
    h = aud.model.feature_extractor.conv_layers[0].register_forward_hook(lambda *_: calls.append(1))
    aud.prefetch_features(x)
    aud(x, attention_mask=torch.ones(2, 8000, dtype=torch.bool))
    aud(x, attention_mask=torch.ones(2, 8000, dtype=torch.bool))
    h.remove()
    assert len(calls) == 1


def test_frozen_cast_cache_changes_no_value_and_follows_weight_updates():
    """install_frozen_cast_cache: under autocast a frozen Linear/Conv gets its lower-precision weight from a cache
    instead of a per-call cast; outputs are bit-identical, trainable modules are left alone, in-place weight edits
    and (un)freezing are picked up."""
    from multimodal_av_model_b200.encoders import install_frozen_cast_cache
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv1d(3, 8, 3, padding=1), torch.nn.PReLU(8), torch.nn.Conv1d(8, 8, 1),
                              torch.nn.Flatten(), torch.nn.Linear(8 * 16, 5))
    for p in list(net[0].parameters()) + list(net[1].parameters()) + list(net[4].parameters()):
        p.requires_grad = False                                       # net[2] stays trainable
    x = torch.randn(4, 3, 16)

    def run():
        with torch.autocast("cpu", dtype=torch.bfloat16):
            return net(x)
    ref = run()
    assert install_frozen_cast_cache(net) == 4
    got = run()
    assert torch.equal(ref, got) and got.dtype == torch.bfloat16
    assert set(net[0]._avctc_cast_cache) == {"weight", "bias"} and not net[2]._avctc_cast_cache
    assert isinstance(net[0].weight, torch.nn.Parameter) and net[0].weight.dtype == torch.float32   # restored after the call
    got.float().sum().backward()
    assert net[2].weight.grad is not None and net[0].weight.grad is None
    with torch.no_grad():
        net[4].weight.mul_(2.0)                                       # in-place edit -> version bump -> cache refreshed
    fresh = run()
    assert not torch.equal(fresh, got)
    net[4].weight.requires_grad = True                                # unfrozen: normal autocast path, same values
    assert torch.equal(run(), fresh)
    assert torch.equal(net(x), net(x)) and net(x).dtype == torch.float32      # no autocast: untouched fp32 path


def test_frozen_cast_cache_on_both_encoders_is_value_neutral():
    from multimodal_av_model_b200.encoders import install_frozen_cast_cache
    torch.manual_seed(0)
    vis = pkg.VisualEncoder(relu_type="prelu").eval()
    for p in vis.parameters():
        p.requires_grad = False
    lips = torch.rand(1, 1, 6, 96, 96)
    aud = _tiny_audio().eval()
    unfreeze_middle_layers(aud.model)
    x = 0.1 * torch.randn(2, 8000)
    m = torch.ones(2, 8000, dtype=torch.bool); m[1, 6000:] = False

    def run():
        with torch.autocast("cpu", dtype=torch.bfloat16), torch.no_grad():
            return vis(lips), aud(x, attention_mask=m)
    v0, (a0, mid0) = run()
    assert install_frozen_cast_cache(vis) > 10 and install_frozen_cast_cache(aud) > 10
    v1, (a1, mid1) = run()
    assert torch.equal(v0, v1) and torch.equal(a0, a1) and torch.equal(mid0, mid1)
