"""GPU: CUDA-graph replay of the frozen encoder segments (graphed.py) changes no value and no side effect:
same outputs as the eager modules, BatchNorm running statistics advance once per call (capture warm-up leaves no trace),
a changed parameter (load_state_dict) is picked up, and a gradient-requiring call takes the eager path."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


def test_visual_encoder_graph_equals_eager_and_bn_stats_advance_once_per_call():
    pkg = _pkg()
    torch.manual_seed(0)
    a = pkg.VisualEncoder().cuda()
    for p in a.parameters():
        p.requires_grad = False
    b = copy.deepcopy(a)
    a.train(); b.train()
    xs = [torch.rand(2, 1, 9, 96, 96, device="cuda") for _ in range(4)]
    outs = {}
    for name, m, flag in (("graph", a, 1), ("eager", b, 0)):
        pkg._lib.set_py_tuning("enc_graphs", flag)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs[name] = [m(x).float().clone() for x in xs]
        finally:
            pkg._lib.set_py_tuning("enc_graphs", 1)
    # a shape is captured the second time it is met: call 1 eager, call 2 capture + replay, calls 3-4 replay
    assert a._graph_seg.captures == 1 and a._graph_seg.replays == 3 and a._graph_seg.eager == 1
    assert not hasattr(b, "_graph_seg") or b._graph_seg.replays == 0
    for y, z in zip(outs["graph"], outs["eager"]):
        assert torch.allclose(y, z, rtol=2e-2, atol=2e-2)
    bn_a, bn_b = a.frontend3D[1], b.frontend3D[1]
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 4
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-3, atol=1e-5)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-3, atol=1e-5)
    la, lb = a.trunk.layer4[1].bn2, b.trunk.layer4[1].bn2
    assert torch.allclose(la.running_mean, lb.running_mean, rtol=2e-2, atol=1e-3)
    # eval mode is a different kernel sequence (BatchNorm uses the running statistics): never the train-mode graph
    a.eval(); b.eval()
    outs_eval = {}
    for name, m, flag in (("graph", a, 1), ("eager", b, 0)):
        pkg._lib.set_py_tuning("enc_graphs", flag)
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                outs_eval[name] = [m(xs[0]).float().clone() for _ in range(3)][-1]
        finally:
            pkg._lib.set_py_tuning("enc_graphs", 1)
    assert a._graph_seg.captures == 2
    assert torch.allclose(outs_eval["graph"], outs_eval["eager"], rtol=2e-2, atol=2e-2)
    assert not torch.allclose(outs_eval["graph"], outs["graph"][0], atol=1e-2)     # batch statistics vs running statistics
    assert int(bn_a.num_batches_tracked) == 4
    a.train(); b.train()
    # a reloaded weight must be seen by the next call (new parameter version -> new capture)
    sd = {k: (v * 0 if k == "frontend3D.0.weight" else v) for k, v in a.state_dict().items()}
    a.load_state_dict(sd)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a(xs[0])                                  # new parameter versions = new signature: eager once, then captured
        y0 = a(xs[0]).float()
    assert a._graph_seg.captures == 3
    assert not torch.allclose(y0, outs["graph"][0], atol=1e-3)


def test_audio_encoder_graphed_segments_equal_eager_in_eval():
    pkg = _pkg()
    from multimodal_av_model_b200.encoders import unfreeze_middle_layers, xlsr_large_config
    cfg = xlsr_large_config(hidden_size=64, num_hidden_layers=10, num_attention_heads=4, intermediate_size=128,
                            conv_dim=(32,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    torch.manual_seed(0)
    aud = pkg.AudioEncoder(freeze=True, config=cfg).cuda()
    unfreeze_middle_layers(aud.model)
    aud.eval()
    x = 0.1 * torch.randn(2, 16000, device="cuda")
    m = torch.ones(2, 16000, dtype=torch.bool, device="cuda"); m[1, 12000:] = False
    res = {}
    for flag in (0, 1):
        pkg._lib.set_py_tuning("enc_graphs", flag)
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                for _ in range(2):                # second sight of a shape captures it
                    res[flag] = [t.float().clone() for t in aud(x.clone(), attention_mask=m)]
        finally:
            pkg._lib.set_py_tuning("enc_graphs", 1)
    for y, z in zip(res[0], res[1]):
        assert torch.allclose(y, z, rtol=2e-2, atol=2e-2)
    segs = [getattr(l, "_avctc_graph_seg", None) for l in aud.model.encoder.layers]
    assert segs[0] is not None and segs[0].replays >= 1                 # frozen layer: replayed
    assert segs[6] is None or segs[6].replays == 0                      # trainable layer: never graphed
    # training: gradients still reach the trainable layers and only the layers below them are replayed
    aud.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        last, mid = aud(x.clone(), attention_mask=m)
    (last.float().sum() + mid.float().sum()).backward()
    assert any(p.grad is not None for n, p in aud.model.named_parameters() if "encoder.layers.7." in n)
    r9 = getattr(aud.model.encoder.layers[9], "_avctc_graph_seg", None)
    assert r9 is None or r9.replays == 0


def test_graph_capture_policy_does_not_thrash_on_varying_shapes():
    """Padded batch shapes of a real data loader vary from batch to batch: a shape met once runs eagerly (no capture),
    and a stream of always-new shapes never captures."""
    pkg = _pkg()
    torch.manual_seed(0)
    v = pkg.VisualEncoder().cuda()
    for p in v.parameters():
        p.requires_grad = False
    v.eval()
    with torch.no_grad():
        for T in range(3, 13):
            v(torch.rand(1, 1, T, 96, 96, device="cuda"))
    assert v._graph_seg.captures == 0 and v._graph_seg.eager == 10 and v._graph_seg.replays == 0
