"""oracle/torch_port.py — the reference's hot path restated on torch CPU ops.  TEST INFRASTRUCTURE ONLY.

The reference is ~150 lines of Python glue over un-vendored PyTorch ops, so "what the reference executes
on the host CPU" is those same ATen kernels.  This file restates the glue (it is NOT a copy: the
per-sample Python loops are replaced by vectorised index arithmetic with identical results) so that
  * tests can check gradients at the real sizes against the arithmetic the reference really runs, and
  * bench.py's cpu_baseline / --impl reference legs can time that arithmetic on the GPU box, where
    /root/reference does not exist.
tests/test_oracle_golden.py::test_torch_port_* pins every function here to the fixtures produced by the
reference itself (tests/golden/, oracle/gen_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

TEMPERATURE = 0.07          # /root/reference/contrastive.py:4-6
WEIGHT_POS_ALIGN = 1.0
WEIGHT_NEG_SUPPRESS = 0.3


def select_pad_resample(audio, mask, t_v):
    """fusion_module.py:40-55 without the Python loop: stable compaction of speech frames to the front,
    zero pad to the batch max, then the same F.interpolate calls."""
    speech = (mask != 0) & (mask != 3)
    lens = speech.sum(1)
    Tp = int(lens.max())
    order = torch.argsort((~speech).to(torch.int8), dim=1, stable=True)[:, :Tp]
    keep = (torch.arange(Tp, device=mask.device)[None, :] < lens[:, None])
    a = torch.gather(audio, 1, order[:, :, None].expand(-1, -1, audio.shape[2])) * keep[:, :, None].to(audio.dtype)
    m = torch.gather(mask, 1, order) * keep.to(mask.dtype)
    if t_v != Tp:
        a = F.interpolate(a.permute(0, 2, 1), size=t_v, mode="linear", align_corners=True).permute(0, 2, 1)
        m = F.interpolate(m.unsqueeze(1).float(), size=t_v, mode="nearest").squeeze(1).long()
    return a, m


class FusionPort(nn.Module):
    """model/fusion_module.py:5-67 (same submodule names -> same state_dict keys)."""

    def __init__(self, visual_dim, audio_dim, fused_dim, num_heads=4):
        super().__init__()
        self.visual_proj = nn.Linear(visual_dim, fused_dim)
        self.audio_proj = nn.Linear(audio_dim, fused_dim)
        self.cross_attn_visual = nn.MultiheadAttention(fused_dim, num_heads, batch_first=True)
        self.cross_attn_audio = nn.MultiheadAttention(fused_dim, num_heads, batch_first=True)
        self.fusion_proj = nn.Linear(fused_dim, fused_dim)
        self.temporal_model = nn.LSTM(fused_dim, fused_dim, num_layers=2, batch_first=True, bidirectional=True)

    def projection(self, visual_feat, audio_feat, mask):
        a_in, m = select_pad_resample(audio_feat, mask, visual_feat.shape[1])
        v = self.visual_proj(visual_feat)
        a = self.audio_proj(a_in)
        a2v, _ = self.cross_attn_audio(query=a, key=v, value=v)
        return self.fusion_proj(a2v), m

    def forward(self, visual_feat, audio_feat, mask):
        fused, m = self.projection(visual_feat, audio_feat, mask)
        out, _ = self.temporal_model(fused)
        return out, (m != 0).sum(1)


class DecoderPort(nn.Module):
    """model/decoder.py:6-35."""

    def __init__(self, input_dim, vocab_size, blank_id=0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, vocab_size))
        self.blank_id = blank_id

    def forward(self, x):
        return F.log_softmax(self.net(x), dim=-1)


def contrastive_loss_with_mask(middle_feat, flat_mask, projection_layer=None):
    """contrastive.py:8-44."""
    B, T, D = middle_feat.shape
    flat = middle_feat.reshape(B * T, D)
    valid = flat_mask != 3
    feat, mk = flat[valid], flat_mask[valid]
    if projection_layer is not None:
        feat = projection_layer(feat)
    feat = F.normalize(feat, dim=1)
    weak, strong, neg = feat[mk == 1], feat[mk == 2], feat[mk == 0]
    total = torch.zeros((), device=middle_feat.device, requires_grad=True)
    for other, wt in ((strong, WEIGHT_POS_ALIGN), (neg, WEIGHT_NEG_SUPPRESS)):
        if weak.shape[0] > 0 and other.shape[0] > 0:
            total = total + wt * (-F.log_softmax(weak @ other.T / TEMPERATURE, dim=1).mean())
    return total


def downsample_mask(mask, t_enc):
    """trainer.py:98-103."""
    return F.interpolate(mask.unsqueeze(1).float(), size=t_enc, mode="nearest").squeeze(1).long()


def simple_beam_search(log_probs, beam_width=5, blank=0):
    """beam_search.py:2-42 with the same per-frame torch.topk and Python-float scores."""
    T = log_probs.shape[0]
    beams = [((), 0.0)]
    for t in range(T):
        vals, ids = torch.topk(log_probs[t], beam_width)
        vals, ids = vals.tolist(), ids.tolist()
        cand = {}
        for seq, score in beams:
            for c, lp in zip(ids, vals):
                key = seq + (c,)
                s = score + lp
                if key not in cand or s > cand[key]:
                    cand[key] = s
        beams = sorted(cand.items(), key=lambda kv: kv[1], reverse=True)[:beam_width]
    out, prev = [], None
    for c in beams[0][0]:
        if c != prev and c != blank:
            out.append(c)
        prev = c
    return out


def hot_path_losses(fusion, decoder, projection_layer, feats, blank, lambda_=0.1):
    """trainer.py:98-119 from encoder features on.  feats: list (one per speaker) of dicts with
    visual[B,Tv,Dv], audio[B,Tenc,Da], middle[B,Tenc,Da], mask[B,N], text[B,L], text_len[B]."""
    crit = nn.CTCLoss(blank=blank, zero_infinity=True)
    ctc, con = 0, 0
    for f in feats:
        t_enc = f["audio"].shape[1]
        mask_ds = downsample_mask(f["mask"], t_enc)
        con = con + contrastive_loss_with_mask(f["middle"], mask_ds.reshape(-1), projection_layer)
        fused, il = fusion(f["visual"], f["audio"], mask_ds)
        lp = decoder(fused)
        ctc = ctc + crit(lp.transpose(0, 1), f["text"], il, f["text_len"])
    return ctc / 2 + lambda_ * con / 2


def _beam_pool_init():
    torch.set_num_threads(1)


def _beam_pool_job(args):
    lp, beam, blank = args                                  # numpy [T,V] (plain pickling; no shared-memory handles)
    return simple_beam_search(torch.from_numpy(lp), beam, blank)


def beam_search_pool(log_probs, beam_width, blank, workers):
    """simple_beam_search over a batch [N,T,V] in a pool of `workers` single-threaded processes (the reference decodes one
    utterance at a time in Python, trainer.py:229-242; a pool over the host cores is the best a CPU box can do with it).
    Returns (seconds of decode wall time excluding pool start-up, token lists)."""
    import multiprocessing as mp
    import time
    lp = log_probs.detach().float().cpu().contiguous()
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_beam_pool_init) as pool:
        pool.map(_beam_pool_job, [(lp[0][:2].numpy().copy(), beam_width, blank)] * workers)   # start-up + imports, untimed
        jobs = [(lp[i].numpy(), beam_width, blank) for i in range(lp.shape[0])]
        t0 = time.perf_counter()
        out = pool.map(_beam_pool_job, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
        dt = time.perf_counter() - t0
    return dt, out
