"""oracle/np_oracle.py — numpy float64 restatement of the reference's fusion / CTC-head / InfoNCE
arithmetic.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's modules are thin Python over un-vendored PyTorch ops; each function below cites the
reference call site it follows and restates the published semantics of the torch op it lands on
(torch 2.11.0: nn.Linear, nn.MultiheadAttention -> F.multi_head_attention_forward
torch/nn/functional.py:6244-6672, nn.LSTM gate order i,f,g,o, F.interpolate index rules pinned in
SURVEY.md §3.5).  Pinned by tests/golden/fusion_*.npz, infonce_*.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np

TEMPERATURE = 0.07          # /root/reference/contrastive.py:4
WEIGHT_POS_ALIGN = 1.0      # /root/reference/contrastive.py:5
WEIGHT_NEG_SUPPRESS = 0.3   # /root/reference/contrastive.py:6


# ----------------------------------------------------------------------------- interpolation
def nearest_src_index(out_size: int, in_size: int) -> np.ndarray:
    """F.interpolate(mode='nearest') source index (fusion_module.py:55, trainer.py:98,102,209,218):
    src = min(floor(dst * float32(in/out)), in-1), evaluated in float32 like ATen."""
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.floor(dst * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def linear_align_corners(x: np.ndarray, out_size: int) -> np.ndarray:
    """F.interpolate(mode='linear', align_corners=True) along axis 1 of x[B,T_in,D]
    (fusion_module.py:51): pos = dst*(in-1)/(out-1); lerp(floor(pos), floor(pos)+1)."""
    B, Tin, D = x.shape
    if out_size == Tin:
        return x.copy()
    scale = (Tin - 1) / (out_size - 1) if out_size > 1 else 0.0
    pos = np.arange(out_size, dtype=np.float64) * scale
    i0 = np.minimum(np.floor(pos).astype(np.int64), Tin - 1)
    i1 = np.minimum(i0 + 1, Tin - 1)
    w1 = (pos - i0)[None, :, None]
    return (1.0 - w1) * x[:, i0, :] + w1 * x[:, i1, :]


def downsample_mask(mask: np.ndarray, t_enc: int) -> np.ndarray:
    """trainer.py:98-103: F.interpolate(mask.float(), size=T_enc, mode='nearest').long()."""
    return mask[:, nearest_src_index(t_enc, mask.shape[1])]


# ----------------------------------------------------------------------------- fusion front half
def select_pad_resample(audio: np.ndarray, mask: np.ndarray, t_v: int):
    """fusion_module.py:40-55.  audio [B,T_a,D], mask [B,T_a] in {0,1,2,3} ->
    (audio [B,T_v,D], mask [B,T_v]).  speech = mask not in {0,3}; per-sample compaction; zero pad
    to the batch max; if T_v != padded length: linear(align_corners) for audio, nearest for mask."""
    B, Ta, D = audio.shape
    speech = (mask != 0) & (mask != 3)
    lens = speech.sum(1)
    Tp = int(lens.max()) if B else 0
    a = np.zeros((B, Tp, D), dtype=np.float64)
    m = np.zeros((B, Tp), dtype=np.int64)
    for i in range(B):
        a[i, :lens[i]] = audio[i][speech[i]]
        m[i, :lens[i]] = mask[i][speech[i]]
    if t_v != Tp:
        a = linear_align_corners(a, t_v)
        m = m[:, nearest_src_index(t_v, Tp)]
    return a, m


def _softmax(x, axis=-1):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def cross_attention(a, v, in_w, in_b, out_w, out_b, num_heads):
    """nn.MultiheadAttention(batch_first=True)(query=a, key=v, value=v) (fusion_module.py:61):
    packed in-projection rows [q;k;v], scaled dot-product over ALL keys (no mask), out_proj."""
    B, T, E = a.shape
    hd = E // num_heads
    q = a @ in_w[:E].T + in_b[:E]
    k = v @ in_w[E:2 * E].T + in_b[E:2 * E]
    val = v @ in_w[2 * E:].T + in_b[2 * E:]
    Tk = v.shape[1]
    q = q.reshape(B, T, num_heads, hd).transpose(0, 2, 1, 3)
    k = k.reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    val = val.reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    p = _softmax((q / np.sqrt(hd)) @ k.transpose(0, 1, 3, 2), axis=-1)
    o = (p @ val).transpose(0, 2, 1, 3).reshape(B, T, E)
    return o @ out_w.T + out_b


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM over ALL T frames (no packing; fusion_module.py:64). Gates i,f,g,o."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H)); c = np.zeros((B, H))
    out = np.zeros((B, T, H))
    xs = x @ w_ih.T + b_ih + b_hh
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = xs[:, t] + h @ w_hh.T
        i = _sigmoid(g[:, :H]); f = _sigmoid(g[:, H:2 * H])
        gg = np.tanh(g[:, 2 * H:3 * H]); o = _sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * np.tanh(c)
        out[:, t] = h
    return out


def bilstm2(x, p, prefix="temporal_model."):
    for layer in range(2):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            outs.append(lstm_direction(
                x, p[f"{prefix}weight_ih_l{layer}{suffix}"], p[f"{prefix}weight_hh_l{layer}{suffix}"],
                p[f"{prefix}bias_ih_l{layer}{suffix}"], p[f"{prefix}bias_hh_l{layer}{suffix}"], rev))
        x = np.concatenate(outs, axis=-1)
    return x


def fusion_forward(p: dict, visual, audio, mask, num_heads=4, upto="lstm"):
    """CrossAttentionFusion.forward (fusion_module.py:29-67) with state_dict `p` (numpy float64).
    Returns (fused_seq [B,T_v,2E], input_lengths [B]).  upto='proj' stops after fusion_proj."""
    visual = np.asarray(visual, dtype=np.float64)
    audio = np.asarray(audio, dtype=np.float64)
    mask = np.asarray(mask, dtype=np.int64)
    t_v = visual.shape[1]
    a_in, m = select_pad_resample(audio, mask, t_v)
    v = visual @ p["visual_proj.weight"].T + p["visual_proj.bias"]          # :57
    a = a_in @ p["audio_proj.weight"].T + p["audio_proj.bias"]              # :58
    a2v = cross_attention(a, v, p["cross_attn_audio.in_proj_weight"], p["cross_attn_audio.in_proj_bias"],
                          p["cross_attn_audio.out_proj.weight"], p["cross_attn_audio.out_proj.bias"],
                          num_heads)                                          # :61
    fused = a2v @ p["fusion_proj.weight"].T + p["fusion_proj.bias"]          # :63
    input_lengths = (m != 0).sum(1).astype(np.int64)                         # :66
    if upto == "proj":
        return fused, input_lengths
    return bilstm2(fused, p), input_lengths                                  # :64


def ctc_head(x, w, b):
    """CTCDecoder.forward without target (decoder.py:24-25,35): log_softmax(x W^T + b)."""
    z = np.asarray(x, dtype=np.float64) @ w.T + b
    z = z - z.max(-1, keepdims=True)
    return z - np.log(np.exp(z).sum(-1, keepdims=True))


# ----------------------------------------------------------------------------- InfoNCE
def contrastive_loss_with_mask(middle, flat_mask, w=None, b=None, want_grad=False):
    """contrastive.py:8-44.  middle [B,T,D]; flat_mask [B*T] in {0,1,2,3}; optional projection
    (w [P,D], b [P]).  loss = 1.0*mean(-log_softmax(Aw.As^T/0.07)) + 0.3*mean(-log_softmax(Aw.An^T/0.07))
    with rows L2-normalised (F.normalize eps 1e-12).  want_grad: also returns analytic
    d loss/d middle, d/dw, d/db."""
    middle = np.asarray(middle, dtype=np.float64)
    B, T, D = middle.shape
    flat = middle.reshape(B * T, D)
    fm = np.asarray(flat_mask, dtype=np.int64).reshape(-1)
    valid = np.nonzero(fm != 3)[0]
    x = flat[valid]
    mk = fm[valid]
    y = x @ w.T + b if w is not None else x
    n = np.sqrt((y * y).sum(1, keepdims=True))
    nc = np.maximum(n, 1e-12)
    z = y / nc
    strong = np.nonzero(mk == 2)[0]
    weak = np.nonzero(mk == 1)[0]
    neg = np.nonzero(mk == 0)[0]
    loss = 0.0
    dz = np.zeros_like(z)
    for other, wt in ((strong, WEIGHT_POS_ALIGN), (neg, WEIGHT_NEG_SUPPRESS)):
        if len(weak) == 0 or len(other) == 0:
            continue
        A = z[weak]; S = z[other]
        sim = A @ S.T / TEMPERATURE
        mx = sim.max(1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(sim - mx).sum(1))
        loss += wt * (lse.mean() - sim.mean())
        if want_grad:
            P = np.exp(sim - lse[:, None])
            dsim = wt * (P / len(weak) - 1.0 / (len(weak) * len(other))) / TEMPERATURE
            np.add.at(dz, weak, dsim @ S)
            np.add.at(dz, other, dsim.T @ A)
    if not want_grad:
        return loss
    dy = (dz - z * (z * dz).sum(1, keepdims=True)) / nc
    dy = np.where(n > 1e-12, dy, dz / 1e-12)
    if w is not None:
        dx = dy @ w
        dw = dy.T @ x
        db = dy.sum(0)
    else:
        dx, dw, db = dy, None, None
    dflat = np.zeros_like(flat)
    dflat[valid] = dx
    return loss, dflat.reshape(B, T, D), dw, db
