"""GPU: the drop-in trainer — hot-path loss against the reference's own step fixture, and an end-to-end
train_epoch / evaluate on tiny random-init encoders."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


class _Enc(torch.nn.Module):
    def forward(self, *a, **k):
        raise AssertionError("encoders are not used by hot_path_loss")


def test_hot_path_loss_matches_reference_step_fixture():
    """trainer.py:98-119 from encoder features on, vs tests/golden/step_cases.npz (reference, fp32 CPU).
    bf16 operands inside the fusion GEMMs -> 2e-2 on the loss, 6e-2 (of the max) on parameter gradients."""
    import os
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer
    z = np.load(os.path.join(GOLDEN, "step_cases.npz"))
    p = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    e = p["fusion_proj.weight"].shape[0]
    fus = pkg.CrossAttentionFusion(p["visual_proj.weight"].shape[1], p["audio_proj.weight"].shape[1], e)
    fus.load_state_dict(p)
    dec = pkg.CTCDecoder(2 * e, z["dec/net.0.weight"].shape[0], blank_id=3)
    dec.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dec/")})
    tr = pkg.MultimodalTrainer(_Enc(), _Enc(), fus, dec, CharTokenizer(30), device="cuda")
    tr.projection_layer = torch.nn.Linear(z["proj/weight"].shape[1], z["proj/weight"].shape[0]).cuda()
    tr.projection_layer.load_state_dict({"weight": torch.from_numpy(z["proj/weight"]), "bias": torch.from_numpy(z["proj/bias"])})
    c = lambda k: torch.from_numpy(z[k]).cuda()
    total, c1, c2, k1, k2 = tr.hot_path_loss([c("spk1/visual"), c("spk2/visual")], [c("spk1/audio"), c("spk2/audio")],
                                             [c("spk1/middle"), c("spk2/middle")], [c("spk1/mask"), c("spk2/mask")],
                                             [c("spk1/text"), c("spk2/text")], [c("spk1/text_len"), c("spk2/text_len")])
    total.backward()
    assert abs(k1.item() - float(z["spk1/contrast"])) < 1e-4 * abs(float(z["spk1/contrast"]))   # fp32 features -> fp32 path
    assert abs(c1.item() - float(z["spk1/ctc"])) < 2e-2 * abs(float(z["spk1/ctc"]))
    assert abs(total.item() - float(z["loss_total"])) < 2e-2 * abs(float(z["loss_total"]))
    for k, prm in fus.named_parameters():
        g = z[f"grad/{k}"]
        if g.size == 0:
            assert prm.grad is None
        else:
            err = (prm.grad.cpu() - torch.from_numpy(g)).abs().max() / (np.abs(g).max() + 1e-12)
            assert err < 6e-2, (k, float(err))


def tiny_models(pkg):
    from multimodal_av_model_b200.encoders import xlsr_large_config
    cfg = xlsr_large_config(hidden_size=64, num_hidden_layers=10, num_attention_heads=4, intermediate_size=128,
                            conv_dim=(32,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4)
    torch.manual_seed(0)
    vis = pkg.VisualEncoder()
    for prm in vis.parameters():
        prm.requires_grad = False
    aud = pkg.AudioEncoder(freeze=True, config=cfg)
    from multimodal_av_model_b200.encoders import unfreeze_middle_layers
    unfreeze_middle_layers(aud.model)
    fus = pkg.CrossAttentionFusion(512, 64, 64)
    dec = pkg.CTCDecoder(128, 800, blank_id=3)
    return vis, aud, fus, dec


def test_train_epoch_and_evaluate_end_to_end():
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_batch
    vis, aud, fus, dec = tiny_models(pkg)
    tr = pkg.MultimodalTrainer(vis, aud, fus, dec, CharTokenizer(800), device="cuda")
    tr.verbose = False
    batches = [make_batch(pairs=2, seconds=1.0, t_v=30, seed=s, l_range=(3, 8)) for s in range(3)]
    w0 = dec.net[0].weight.detach().clone()
    l0 = tr.train_epoch(batches)
    l1 = tr.train_epoch(batches)
    assert np.isfinite(l0) and np.isfinite(l1) and l0 > 0
    assert tr.last_epoch_steps == 3 and tr.loss_log_count == 3        # every step ran; its loss was logged to the host
    assert float(tr.loss_log[:3].mean()) == pytest.approx(l1, rel=1e-5)
    assert not torch.equal(w0, dec.net[0].weight.detach())            # the optimiser moved the CTC head
    assert fus.cross_attn_visual.in_proj_weight.grad is None         # never used, like the reference
    assert any(p.grad is not None for n, p in aud.model.named_parameters() if "encoder.layers.7." in n)
    loss, wer = tr.evaluate(batches[:2])
    assert np.isfinite(loss) and 0.0 <= wer
    assert tr.ctc_decode([5, 5, 3, 5, 6, 3, 3, 6]) == [5, 6]          # blank does not reset prev (trainer.py:168-177)


def test_train_epoch_prefetch_changes_no_value():
    """train_epoch stages batch i+1 (side-stream H2D) while step i runs; the losses are those of the plain loop."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_batch
    batches = [make_batch(pairs=2, seconds=1.0, t_v=30, seed=s, l_range=(3, 8), pin=(s % 2 == 0)) for s in range(4)]
    logs = {}
    for prefetch in (False, True):
        vis, aud, fus, dec = tiny_models(pkg)
        tr = pkg.MultimodalTrainer(vis, aud, fus, dec, CharTokenizer(800), device="cuda")
        tr.verbose = False
        tr.prefetch_batches = prefetch
        torch.manual_seed(11)
        np.random.seed(11)                      # transformers draws the SpecAugment spans from numpy's global RNG
        avg = tr.train_epoch(batches)
        assert tr.last_epoch_steps == 4
        logs[prefetch] = (avg, tr.loss_log[:4].clone())
    assert logs[True][0] == pytest.approx(logs[False][0], rel=2e-3)
    assert torch.allclose(logs[True][1], logs[False][1], rtol=5e-3)     # fp32 atomics order in split-K weight gradients


def test_train_epoch_keeps_going_after_a_bad_batch(capsys):
    """trainer.py:162-164: a batch that raises is reported and skipped, also when it fails while being staged."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_batch
    vis, aud, fus, dec = tiny_models(pkg)
    tr = pkg.MultimodalTrainer(vis, aud, fus, dec, CharTokenizer(800), device="cuda")
    tr.verbose = False
    good = make_batch(pairs=2, seconds=1.0, t_v=30, seed=0, l_range=(3, 8))
    bad = {k: v for k, v in good.items() if k != "lip2"}             # KeyError inside stage()
    loss = tr.train_epoch([good, bad, good])
    assert tr.last_epoch_steps == 2 and np.isfinite(loss)
    assert "Error at batch 1" in capsys.readouterr().out


def test_wer_matches_jiwer_definition():
    from multimodal_av_model_b200.trainer import word_error_rate
    assert word_error_rate(["a b c", "d e"], ["a x c", "d e f"]) == pytest.approx(2 / 5)
    assert word_error_rate(["a b"], ["a b"]) == 0.0


def test_batching_both_speakers_through_the_bilstm_changes_no_value():
    """hot_path_loss with batch_speakers (one BiLSTM / CTC-head pass over 2B sequences) equals the per-speaker
    order of the reference (trainer.py:110-117): same losses, same parameter and feature gradients."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_features
    torch.manual_seed(0)
    fus = pkg.CrossAttentionFusion(512, 1024, 512)
    dec = pkg.CTCDecoder(1024, 800, blank_id=3)
    tr = pkg.MultimodalTrainer(_Enc(), _Enc(), fus, dec, CharTokenizer(800), device="cuda")
    f = make_features(pairs=3, t_v=40, t_enc=99, seed=5, n_samples=32000, dtype=torch.bfloat16)
    out = {}
    for mode in (False, True):
        tr.batch_speakers = mode
        fd = {k: [t.cuda() for t in v] for k, v in f.items()}
        for k in ("audio", "middle"):
            fd[k] = [t.requires_grad_() for t in fd[k]]
        for m in (fus, dec):
            m.zero_grad(set_to_none=True)
        if tr.projection_layer is not None:
            tr.projection_layer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            parts = tr.hot_path_loss(fd["visual"], fd["audio"], fd["middle"], fd["masks"], fd["texts"], fd["lens"])
        parts[0].backward()
        out[mode] = ([float(x) for x in parts], {n: p.grad.clone() for n, p in list(fus.named_parameters()) +
                                                 list(dec.named_parameters()) if p.grad is not None},
                     [t.grad.clone() for t in fd["audio"]])
    a, b = out[False], out[True]
    assert np.allclose(a[0], b[0], rtol=1e-5, atol=1e-6), (a[0], b[0])
    assert a[1].keys() == b[1].keys()
    for n in a[1]:       # split-K / atomics order may differ between a B and a 2B launch: fp32 accumulation noise only
        scale = a[1][n].abs().max() + 1e-12
        assert (a[1][n] - b[1][n]).abs().max() <= 2e-3 * scale, n
    for x, y in zip(a[2], b[2]):
        assert (x.float() - y.float()).abs().max() <= 2e-2 * (x.float().abs().max() + 1e-12)


def test_contrastive_on_side_stream_changes_no_value():
    """hot_path_loss runs the two InfoNCE losses on a side stream under the BiLSTM kernels (overlap_contrastive): same
    losses and gradients as the single-stream order, over several back-to-back steps (allocator reuse across streams)."""
    pkg = _pkg()
    from multimodal_av_model_b200.synthetic import CharTokenizer, make_features
    torch.manual_seed(0)
    fus = pkg.CrossAttentionFusion(512, 1024, 512)
    dec = pkg.CTCDecoder(1024, 800, blank_id=3)
    tr = pkg.MultimodalTrainer(_Enc(), _Enc(), fus, dec, CharTokenizer(800), device="cuda")
    f = make_features(pairs=4, t_v=60, t_enc=99, seed=9, n_samples=32000, dtype=torch.bfloat16)
    out = {}
    for mode in (False, True):
        tr.overlap_contrastive = mode
        runs = []
        for rep in range(3):
            fd = {k: [t.cuda() for t in v] for k, v in f.items()}
            for k in ("audio", "middle"):
                fd[k] = [t.requires_grad_() for t in fd[k]]
            for m in (fus, dec):
                m.zero_grad(set_to_none=True)
            if tr.projection_layer is not None:
                tr.projection_layer.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                parts = tr.hot_path_loss(fd["visual"], fd["audio"], fd["middle"], fd["masks"], fd["texts"], fd["lens"])
            parts[0].backward()
            torch.cuda.synchronize()
            runs.append(([float(x.detach()) for x in parts], [t.grad.clone() for t in fd["middle"]],
                         tr.projection_layer.weight.grad.clone()))
        out[mode] = runs
    for a, b in zip(out[False], out[True]):
        assert np.allclose(a[0], b[0], rtol=1e-6, atol=1e-7), (a[0], b[0])
        for x, y in zip(a[1], b[1]):
            assert torch.equal(x, y)                                  # the InfoNCE backward is bit-reproducible
        assert (a[2] - b[2]).abs().max() <= 1e-5 * (a[2].abs().max() + 1e-12)      # split-K atomics order only
