"""CrossAttentionFusion — drop-in for /root/reference/model/fusion_module.py:5-67 on sm_100a kernels.

Same constructor, forward signature, return values and state_dict keys as the reference (its
submodules are kept as parameter containers, so default initialisation consumes the RNG identically):

    CrossAttentionFusion(visual_dim, audio_dim, fused_dim, num_heads=4)
    forward(visual_feat[B,T_v,D_v], audio_feat[B,T_a,D_a], mask[B,T_a] int64) -> (fused_seq[B,T_v,2E], input_lengths[B])

What runs where:
  * speech-frame select + pad + linear/nearest resample + input_lengths (fusion_module.py:40-55,66):
    avctc_resample_* (two launches, no host sync; the reference does B Python iterations and B .item()s)
  * visual_proj, audio_proj, the MultiheadAttention in/out projections, Q.K^T, P.V, fusion_proj
    (fusion_module.py:57-63) forward AND backward: avctc_gemm_bf16 = tcgen05.mma + TMEM + TMA, bf16 operands,
    fp32 accumulation, transposed operands read in place (no transpose copies); softmax in fp32
  * the attention core (scores -> softmax -> P.V and its backward): ONE tcgen05 kernel per direction, scores in tensor
    memory only (csrc/attention.cu; T <= 192, head_dim 128 — other shapes run GEMM + softmax + GEMM)
  * temporal_model (2-layer BiLSTM, fusion_module.py:64): persistent sm_100a kernels (csrc/lstm.cu) for hidden 256/512;
    other shapes use torch.nn.LSTM / cuDNN
`cross_attn_visual` exists but is never used, exactly like the reference (its parameters get no gradient).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from .gemm import gemm, operand

_BF16 = torch.bfloat16


def _bf16(t):
    return t if t.dtype == _BF16 else t.to(_BF16)


def _f32c(t):
    """fp32 contiguous view of a parameter without touching the dispatcher when it already is one."""
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


_WS_BYTES = {}


def _ws_bytes(fn_name, *dims):
    key = (fn_name,) + dims
    v = _WS_BYTES.get(key)
    if v is None:
        v = _WS_BYTES[key] = int(getattr(_lib.lib(), fn_name)(*dims))
    return v


def _linear(x, w, b, out_dtype=_BF16):
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    return gemm(operand(x), operand(w), M, N, K, out, bias=b, bias_mode=1)


def _dgrad(dy, w):
    """dx[M,K] = dy[M,N] . W[N,K]   (W read in place as the MN-major operand)."""
    M, N = dy.shape
    K = w.shape[1]
    out = torch.empty((M, K), dtype=_BF16, device=dy.device)
    return gemm(operand(dy), operand(w, "mn"), M, K, N, out)


def _wgrad(dy, x, out):
    """out[N,K] (fp32) = dy[M,N]^T . x[M,K]   (both operands read in place as MN-major)."""
    M, N = dy.shape
    K = x.shape[1]
    return gemm(operand(dy, "mn"), operand(x, "mn"), N, K, M, out)


def _colsum(x, out=None):
    M, N = x.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x.device)
    with _lib.device_guard(x.device):
        _lib.check(_lib.lib().avctc_colsum(x.data_ptr(), _lib.dtype_enum(x), M, N, x.stride(0), out.data_ptr(), 0,
                                           _lib.stream_ptr(x.device)), "avctc_colsum")
    return out


class _FusionCoreFnPy(torch.autograd.Function):
    """(per-kernel Python orchestration; used when a head is not a multiple of 64 columns wide)
    resample -> visual/audio projections -> cross attention (audio queries, visual keys/values) -> fusion_proj."""

    @staticmethod
    def forward(ctx, visual, audio, mask, num_heads, w_vp, b_vp, w_ap, b_ap, w_in, b_in, w_o, b_o, w_f, b_f):
        _lib.require_cuda(visual, "visual_feat")
        _lib.require_cuda(audio, "audio_feat")
        if mask is None:
            raise RuntimeError("mask is required (the reference indexes it unconditionally, fusion_module.py:44)")
        dev = visual.device
        B, T, Dv = visual.shape
        _, Ta, Da = audio.shape
        E = w_f.shape[0]
        H = int(num_heads)
        hd = E // H
        if E % H or E % 8 or Dv % 8 or Da % 8:
            raise RuntimeError("fused_dim, visual_dim and audio_dim must be multiples of 8 (TMA row alignment)")
        M = B * T
        L = _lib.lib()
        st = _lib.stream_ptr(dev)
        audio_c = audio.detach()
        if audio_c.dtype not in (torch.float32, _BF16):
            audio_c = audio_c.float()
        audio_c = audio_c.contiguous()
        mask_c = mask.to(device=dev, dtype=torch.long).contiguous()
        xa = torch.empty((M, Da), dtype=_BF16, device=dev)
        mask_out = torch.empty((B, T), dtype=torch.long, device=dev)
        input_lengths = torch.empty(B, dtype=torch.long, device=dev)
        rs_bytes = int(L.avctc_resample_workspace_bytes(B, Ta))
        rs_ws = torch.empty(rs_bytes, dtype=torch.uint8, device=dev)
        with _lib.device_guard(dev):
            _lib.check(L.avctc_resample_forward(audio_c.data_ptr(), _lib.dtype_enum(audio_c), mask_c.data_ptr(), B, Ta, Da,
                                                T, xa.data_ptr(), mask_out.data_ptr(), input_lengths.data_ptr(),
                                                rs_ws.data_ptr(), rs_bytes, st), "avctc_resample_forward")
        xv = _bf16(visual.detach().reshape(M, Dv)).contiguous()
        wb = [_bf16(w.detach()).contiguous() for w in (w_vp, w_ap, w_in, w_o, w_f)]
        bs = [b.detach().float().contiguous() for b in (b_vp, b_ap, b_in, b_o, b_f)]
        # The attention GEMMs address head h as 64-wide K blocks starting at column h*hd, so a head must own a
        # multiple of 64 columns.  Other head sizes (toy configs) run on zero-padded in/out projection weights:
        # the padded q/k/v columns are exactly 0, which changes no product.
        hdp = (hd + 63) // 64 * 64
        Ea = H * hdp
        pad_idx = None
        if hdp != hd:
            pad_idx = (torch.arange(H, device=dev)[:, None] * hdp + torch.arange(hd, device=dev)[None, :]).reshape(-1)
            w_in_p = torch.zeros((3 * Ea, E), dtype=_BF16, device=dev)
            b_in_p = torch.zeros(3 * Ea, dtype=torch.float32, device=dev)
            for part in range(3):
                w_in_p[pad_idx + part * Ea] = wb[2][part * E:(part + 1) * E]
                b_in_p[pad_idx + part * Ea] = bs[2][part * E:(part + 1) * E]
            w_o_p = torch.zeros((E, Ea), dtype=_BF16, device=dev)
            w_o_p[:, pad_idx] = wb[3]
            wb[2], bs[2], wb[3] = w_in_p, b_in_p, w_o_p
        Eo, E, hd = E, Ea, hdp          # from here on E / hd are the (possibly padded) attention widths
        v = _linear(xv, wb[0], bs[0])
        a = _linear(xa, wb[1], bs[1])
        q = _linear(a, wb[2][:E], bs[2][:E])
        kv = _linear(v, wb[2][E:], bs[2][E:])                     # [M, 2E]: keys | values
        Tp = (T + 7) // 8 * 8
        BH = B * H
        alpha = float(Eo // H) ** -0.5
        S = torch.empty((BH, T, Tp), dtype=torch.float32, device=dev)
        gemm(operand(q, k_inner=hd, r_outer=T), operand(kv, k_inner=hd, r_outer=T), T, T, hd, S, batch=BH,
             inner_count=H, ldc=Tp, c_outer=H * T * Tp, c_inner=T * Tp, alpha=alpha)
        P = torch.empty((BH, T, Tp), dtype=_BF16, device=dev)
        with _lib.device_guard(dev):
            _lib.check(L.avctc_softmax_forward(S.data_ptr(), P.data_ptr(), BH * T, T, Tp, st), "avctc_softmax_forward")
        o = torch.empty((M, E), dtype=_BF16, device=dev)
        vv = kv[:, E:]
        gemm(operand(P, z_outer=H, z_inner=1, kdim=T), operand(vv, "mn", k_outer=T, r_inner=hd), T, hd, T, o, batch=BH,
             inner_count=H, ldc=E, c_outer=T * E, c_inner=hd)
        ao = _linear(o, wb[3], bs[3])
        f = _linear(ao, wb[4], bs[4], out_dtype=torch.float32)
        ctx.save_for_backward(xv, xa, v, a, q, kv, P, o, ao, rs_ws, *wb)
        ctx.dims = (B, T, Ta, Dv, Da, Eo, E, H, hd, Tp, alpha, audio.dtype, visual.dtype)
        ctx.pad_idx = pad_idx
        ctx.mark_non_differentiable(mask_out, input_lengths)
        return f.view(B, T, Eo), mask_out, input_lengths

    @staticmethod
    def backward(ctx, df, _dm, _dl):
        xv, xa, v, a, q, kv, P, o, ao, rs_ws, wb_vp, wb_ap, wb_in, wb_o, wb_f = ctx.saved_tensors
        B, T, Ta, Dv, Da, Eo, E, H, hd, Tp, alpha, audio_dtype, visual_dtype = ctx.dims   # E, hd: padded widths
        pad_idx = ctx.pad_idx
        dev = df.device
        M = B * T
        BH = B * H
        L = _lib.lib()
        st = _lib.stream_ptr(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        dfb = _bf16(df.reshape(M, Eo)).contiguous()
        # fusion_proj
        g_wf = _wgrad(dfb, ao, torch.empty((Eo, Eo), **f32)); g_bf = _colsum(dfb)
        dao = _dgrad(dfb, wb_f)
        # out_proj
        g_wo = _wgrad(dao, o, torch.empty((Eo, E), **f32)); g_bo = _colsum(dao)
        do = _dgrad(dao, wb_o)
        # attention core
        vv = kv[:, E:]
        dP = torch.empty((BH, T, Tp), **f32)
        gemm(operand(do, k_inner=hd, r_outer=T), operand(vv, k_inner=hd, r_outer=T), T, T, hd, dP, batch=BH,
             inner_count=H, ldc=Tp, c_outer=H * T * Tp, c_inner=T * Tp)
        dS = torch.empty((BH, T, Tp), dtype=_BF16, device=dev)
        with _lib.device_guard(dev):
            _lib.check(L.avctc_softmax_backward(P.data_ptr(), dP.data_ptr(), dS.data_ptr(), BH * T, T, Tp, st),
                       "avctc_softmax_backward")
        dq = torch.empty((M, E), dtype=_BF16, device=dev)
        gemm(operand(dS, z_outer=H, z_inner=1, kdim=T), operand(kv[:, :E], "mn", k_outer=T, r_inner=hd), T, hd, T, dq,
             batch=BH, inner_count=H, ldc=E, c_outer=T * E, c_inner=hd, alpha=alpha)
        dkv = torch.empty((M, 2 * E), dtype=_BF16, device=dev)
        gemm(operand(dS, "mn", z_outer=H, z_inner=1, rows=T, kdim=T), operand(q, "mn", k_outer=T, r_inner=hd), T, hd, T,
             dkv[:, :E], batch=BH, inner_count=H, ldc=2 * E, c_outer=T * 2 * E, c_inner=hd, alpha=alpha)
        gemm(operand(P, "mn", z_outer=H, z_inner=1, rows=T, kdim=T), operand(do, "mn", k_outer=T, r_inner=hd), T, hd, T,
             dkv[:, E:], batch=BH, inner_count=H, ldc=2 * E, c_outer=T * 2 * E, c_inner=hd)
        # in_proj (rows [0,E) = query projection of a; rows [E,3E) = key|value projections of v)
        g_win = torch.empty((3 * E, Eo), **f32)
        g_bin = torch.empty(3 * E, **f32)
        _wgrad(dq, a, g_win[:E]); _colsum(dq, g_bin[:E])
        _wgrad(dkv, v, g_win[E:]); _colsum(dkv, g_bin[E:])
        da = _dgrad(dq, wb_in[:E])
        dv = _dgrad(dkv, wb_in[E:])
        if pad_idx is not None:          # drop the zero-padded head columns again
            rows = torch.cat([pad_idx + part * E for part in range(3)])
            g_win, g_bin, g_wo = g_win[rows].contiguous(), g_bin[rows].contiguous(), g_wo[:, pad_idx].contiguous()
        # audio_proj / visual_proj
        g_wap = _wgrad(da, xa, torch.empty((Eo, Da), **f32)); g_bap = _colsum(da)
        g_wvp = _wgrad(dv, xv, torch.empty((Eo, Dv), **f32)); g_bvp = _colsum(dv)
        d_visual = d_audio = None
        if ctx.needs_input_grad[0]:
            d_visual = _dgrad(dv, wb_vp).view(B, T, Dv).to(visual_dtype)
        if ctx.needs_input_grad[1]:
            dxa = _dgrad(da, wb_ap)
            out_dtype = audio_dtype if audio_dtype in (torch.float32, _BF16) else torch.float32
            d_audio = torch.empty((B, Ta, Da), dtype=out_dtype, device=dev)
            with _lib.device_guard(dev):
                _lib.check(L.avctc_resample_backward(dxa.data_ptr(), B, Ta, Da, T, rs_ws.data_ptr(), d_audio.data_ptr(),
                                                     _lib.dtype_enum(d_audio), st), "avctc_resample_backward")
            d_audio = d_audio.to(audio_dtype)
        return (d_visual, d_audio, None, None, g_wvp, g_bvp, g_wap, g_bap, g_win, g_bin, g_wo, g_bo, g_wf, g_bf)


def _weights_key(params):
    return tuple((p.data_ptr(), p._version) for p in params)


class _FusionCoreFn(torch.autograd.Function):
    """Same computation through avctc_fusion_forward / avctc_fusion_backward: ONE host call enqueues every kernel of the
    step (resample, grouped projection GEMMs, the fused tcgen05 attention kernel forward; grouped dgrad/wgrad GEMMs, the
    fused attention backward, bias column sums backward).  `cache` is the module's dict holding the persistent bf16
    copies of the five weight matrices and the (data_ptr, version) key they were made from."""

    @staticmethod
    def forward(ctx, visual, audio, mask, num_heads, cache, w_vp, b_vp, w_ap, b_ap, w_in, b_in, w_o, b_o, w_f, b_f):
        _lib.require_cuda(visual, "visual_feat")
        _lib.require_cuda(audio, "audio_feat")
        if mask is None:
            raise RuntimeError("mask is required (the reference indexes it unconditionally, fusion_module.py:44)")
        dev = visual.device
        B, T, Dv = visual.shape
        _, Ta, Da = audio.shape
        E = w_f.shape[0]
        H = int(num_heads)
        L = _lib.lib()
        st = _lib.stream_ptr(dev)
        audio_c = audio.detach()
        if audio_c.dtype not in (torch.float32, _BF16):
            audio_c = audio_c.float()
        audio_c = audio_c.contiguous()
        mask_c = mask.to(device=dev, dtype=torch.long).contiguous()
        xv = _bf16(visual.detach().reshape(B * T, Dv)).contiguous()
        ws = [_f32c(t) for t in (w_vp, b_vp, w_ap, b_ap, w_in, b_in, w_o, b_o, w_f, b_f)]
        mats = (w_vp, w_ap, w_in, w_o, w_f)
        saved_bytes = _ws_bytes("avctc_fusion_workspace_bytes", B, T, Ta, Dv, Da, E, H, 0)
        scratch_bytes = _ws_bytes("avctc_fusion_workspace_bytes", B, T, Ta, Dv, Da, E, H, 1)
        w_bytes = _ws_bytes("avctc_fusion_workspace_bytes", B, T, Ta, Dv, Da, E, H, 3)
        if saved_bytes == 0:
            raise RuntimeError("fusion dims not supported by the fused path")
        # bf16 weight copies: one persistent buffer per module and device, re-cast only when a parameter changed
        key = _weights_key(mats)
        wbuf = cache.get("buf")
        refresh = wbuf is None or wbuf.device != dev or wbuf.numel() != w_bytes or cache.get("key") != key
        if wbuf is None or wbuf.device != dev or wbuf.numel() != w_bytes:
            wbuf = cache["buf"] = torch.empty(w_bytes, dtype=torch.uint8, device=dev)
        cache["key"] = key
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        # nn.Linear under autocast returns the autocast dtype; plain fp32 modules return fp32
        out_dtype = _BF16 if (visual.dtype == _BF16 or torch.is_autocast_enabled()) else torch.float32
        out = torch.empty((B, T, E), dtype=out_dtype, device=dev)
        mask_out = torch.empty((B, T), dtype=torch.long, device=dev)
        input_lengths = torch.empty(B, dtype=torch.long, device=dev)
        with _lib.device_guard(dev):
            _lib.check(L.avctc_fusion_forward(xv.data_ptr(), audio_c.data_ptr(), _lib.dtype_enum(audio_c), mask_c.data_ptr(),
                                              *[t.data_ptr() for t in ws], B, T, Ta, Dv, Da, E, H, out.data_ptr(),
                                              _lib.dtype_enum(out), mask_out.data_ptr(), input_lengths.data_ptr(),
                                              wbuf.data_ptr(), w_bytes, int(refresh), saved.data_ptr(), saved_bytes,
                                              scratch.data_ptr(), scratch_bytes, st), "avctc_fusion_forward")
        ctx.save_for_backward(xv, saved)
        ctx.dims = (B, T, Ta, Dv, Da, E, H, saved_bytes, w_bytes, audio.dtype, visual.dtype)
        ctx.weights = (mats, key, wbuf)
        ctx.mark_non_differentiable(mask_out, input_lengths)
        return out, mask_out, input_lengths

    @staticmethod
    def backward(ctx, df, _dm, _dl):
        xv, saved = ctx.saved_tensors
        B, T, Ta, Dv, Da, E, H, saved_bytes, w_bytes, audio_dtype, visual_dtype = ctx.dims
        mats, key, wbuf = ctx.weights
        if _weights_key(mats) != key:        # what autograd's saved-tensor version check says for nn.Linear's weight
            raise RuntimeError("one of the variables needed for gradient computation has been modified by an inplace "
                               "operation: a CrossAttentionFusion weight changed between forward and backward")
        dev = df.device
        L = _lib.lib()
        dfc = df.detach()
        if dfc.dtype not in (torch.float32, _BF16):
            dfc = dfc.float()
        dfc = dfc.contiguous()
        shapes = ((E, Dv), (E,), (E, Da), (E,), (3 * E, E), (3 * E,), (E, E), (E,), (E, E), (E,))
        sizes = [(int(torch.Size(shp).numel()) + 63) // 64 * 64 for shp in shapes]       # 256-byte aligned views
        gbuf = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)                   # ONE fill for all ten gradients
        g, off = [], 0
        for shp, n in zip(shapes, sizes):
            g.append(gbuf[off:off + int(torch.Size(shp).numel())].view(shp))
            off += n
        scratch_bytes = _ws_bytes("avctc_fusion_workspace_bytes", B, T, Ta, Dv, Da, E, H, 2)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        d_visual = torch.empty((B, T, Dv), dtype=_BF16, device=dev) if ctx.needs_input_grad[0] else None
        d_audio = None
        if ctx.needs_input_grad[1]:
            d_audio = torch.empty((B, Ta, Da), dtype=audio_dtype if audio_dtype in (torch.float32, _BF16) else torch.float32,
                                  device=dev)
        with _lib.device_guard(dev):
            _lib.check(L.avctc_fusion_backward(dfc.data_ptr(), _lib.dtype_enum(dfc), xv.data_ptr(), B, T, Ta, Dv, Da, E, H,
                                               *[t.data_ptr() for t in g],
                                               d_visual.data_ptr() if d_visual is not None else None,
                                               d_audio.data_ptr() if d_audio is not None else None,
                                               _lib.dtype_enum(d_audio) if d_audio is not None else 0,
                                               wbuf.data_ptr(), w_bytes, saved.data_ptr(), saved_bytes, scratch.data_ptr(),
                                               scratch_bytes, 1, _lib.stream_ptr(dev)), "avctc_fusion_backward")
        if d_visual is not None:
            d_visual = d_visual.to(visual_dtype)
        if d_audio is not None:
            d_audio = d_audio.to(audio_dtype)
        return (d_visual, d_audio, None, None, None, *g)


class _BiLSTMFn(torch.autograd.Function):
    """temporal_model (2-layer bidirectional LSTM, zero initial state, all padded frames — fusion_module.py:21-27,64)
    on the persistent sm_100a kernels of csrc/lstm.cu.  `weights` are nn.LSTM's 16 flat parameters in their own order."""

    @staticmethod
    def forward(ctx, x, *weights):
        _lib.require_cuda(x, "lstm input")
        dev = x.device
        B, T, In = x.shape
        H = weights[1].shape[1]
        L = _lib.lib()
        xb = _bf16(x.detach()).contiguous()
        ws = [_f32c(w) for w in weights]
        need_grad = any(ctx.needs_input_grad)
        saved_bytes = _ws_bytes("avctc_bilstm_workspace_bytes", B, T, In, H, 0)
        scratch_bytes = _ws_bytes("avctc_bilstm_workspace_bytes", B, T, In, H, 1)
        if saved_bytes == 0:
            raise RuntimeError("LSTM shape not supported by the sm_100a kernels")
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        y = torch.empty((B, T, 2 * H), dtype=_BF16, device=dev)
        ptrs = (ctypes.c_void_p * 16)(*[w.data_ptr() for w in ws])
        with _lib.device_guard(dev):
            _lib.check(L.avctc_bilstm_forward(xb.data_ptr(), B, T, In, H, ptrs, y.data_ptr(), saved.data_ptr(), saved_bytes,
                                              scratch.data_ptr(), scratch_bytes, int(need_grad), _lib.stream_ptr(dev)),
                       "avctc_bilstm_forward")
        ctx.save_for_backward(xb, saved)
        ctx.meta = (B, T, In, H, saved_bytes, x.dtype, [tuple(w.shape) for w in weights])
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, saved = ctx.saved_tensors
        B, T, In, H, saved_bytes, xdtype, shapes = ctx.meta
        dev = dy.device
        L = _lib.lib()
        dyb = _bf16(dy.detach()).contiguous()
        grads = [torch.empty(shp, dtype=torch.float32, device=dev) for shp in shapes]
        scratch_bytes = _ws_bytes("avctc_bilstm_workspace_bytes", B, T, In, H, 2)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        dx = torch.empty((B, T, In), dtype=_BF16, device=dev) if ctx.needs_input_grad[0] else None
        ptrs = (ctypes.c_void_p * 16)(*[g.data_ptr() for g in grads])
        with _lib.device_guard(dev):
            _lib.check(L.avctc_bilstm_backward(dyb.data_ptr(), xb.data_ptr(), B, T, In, H, ptrs,
                                               dx.data_ptr() if dx is not None else None, saved.data_ptr(), saved_bytes,
                                               scratch.data_ptr(), scratch_bytes, _lib.stream_ptr(dev)),
                       "avctc_bilstm_backward")
        if dx is not None:
            dx = dx.to(xdtype)
        return (dx, *grads)


class CrossAttentionFusion(nn.Module):
    def __init__(self, visual_dim, audio_dim, fused_dim, num_heads=4):
        super().__init__()
        # same construction order as the reference (fusion_module.py:10-27): identical RNG use and state_dict keys
        self.visual_proj = nn.Linear(visual_dim, fused_dim)
        self.audio_proj = nn.Linear(audio_dim, fused_dim)
        self.cross_attn_visual = nn.MultiheadAttention(embed_dim=fused_dim, num_heads=num_heads, batch_first=True)
        self.cross_attn_audio = nn.MultiheadAttention(embed_dim=fused_dim, num_heads=num_heads, batch_first=True)
        self.fusion_proj = nn.Linear(fused_dim, fused_dim)
        self.temporal_model = nn.LSTM(input_size=fused_dim, hidden_size=fused_dim, num_layers=2, batch_first=True,
                                      bidirectional=True)
        self.num_heads = num_heads
        self._wcache = {}            # persistent bf16 copies of the weight matrices (see _FusionCoreFn)

    def never_used_parameters(self):
        """cross_attn_visual is constructed (fusion_module.py:14) and never called (:61 uses cross_attn_audio only): its
        parameters are outside every graph.  The data-parallel reducer leaves them out of its buckets."""
        return list(self.cross_attn_visual.parameters())

    def fused_projection(self, visual_feat, audio_feat, mask):
        """Everything up to and including fusion_proj (fusion_module.py:40-63): (fused[B,T,E] fp32, mask[B,T], lengths)."""
        at = self.cross_attn_audio
        E = self.fusion_proj.weight.shape[0]
        fused = (E % self.num_heads == 0 and (E // self.num_heads) % 64 == 0 and visual_feat.shape[-1] % 8 == 0
                 and audio_feat.shape[-1] % 8 == 0)
        params = (self.visual_proj.weight, self.visual_proj.bias, self.audio_proj.weight, self.audio_proj.bias,
                  at.in_proj_weight, at.in_proj_bias, at.out_proj.weight, at.out_proj.bias,
                  self.fusion_proj.weight, self.fusion_proj.bias)
        if fused:
            return _FusionCoreFn.apply(visual_feat, audio_feat, mask, self.num_heads, self._wcache, *params)
        return _FusionCoreFnPy.apply(visual_feat, audio_feat, mask, self.num_heads, *params)

    def forward(self, visual_feat, audio_feat, mask=None):
        fused, _mask_rs, input_lengths = self.fused_projection(visual_feat, audio_feat, mask)
        return self.temporal(fused), input_lengths

    def forward_pair(self, visual_feats, audio_feats, masks):
        """Both speakers of a mixed pair: forward(visual_feats[s], audio_feats[s], masks[s]) for s = 0, 1 with ONE pass
        of the recurrent model over the 2B concatenated sequences.  The projections / attention stay per speaker (the
        reference resamples the audio stream to the longest speech segment of ITS batch, fusion_module.py:47-55, so
        the two speakers must not share a batch there); the BiLSTM treats every sequence independently, so running
        it once over both halves the number of sequential time steps and changes no value.
        Returns ((fused_seq_0, fused_seq_1), (input_lengths_0, input_lengths_1))."""
        f, lens = [], []
        for s in range(2):
            fs, _m, il = self.fused_projection(visual_feats[s], audio_feats[s], masks[s])
            f.append(fs); lens.append(il)
        if f[0].shape[1:] != f[1].shape[1:] or f[0].shape[0] + f[1].shape[0] > 64:
            return (self.temporal(f[0]), self.temporal(f[1])), tuple(lens)
        y = self.temporal(torch.cat(f, dim=0))
        b0 = f[0].shape[0]
        return (y[:b0], y[b0:]), tuple(lens)

    def temporal(self, fused):
        """temporal_model over all padded frames (fusion_module.py:64).  The reference's configuration (hidden 512, also
        256; up to 64 sequences) runs on the persistent kernels of csrc/lstm.cu; other shapes use nn.LSTM (cuDNN) and
        say so once."""
        lstm = self.temporal_model
        B, _, In = fused.shape
        max_b = 64 if lstm.hidden_size == 512 else 128
        if (fused.is_cuda and lstm.hidden_size in (256, 512) and B <= max_b and In % 8 == 0 and lstm.num_layers == 2
                and lstm.bidirectional and _lib.tuning_enabled("lstm_custom")):
            y = _BiLSTMFn.apply(fused, *lstm._flat_weights)
            if fused.dtype == torch.float32 and not torch.is_autocast_enabled():
                y = y.float()
            return y
        if fused.is_cuda and _lib.tuning_enabled("lstm_custom") and not getattr(self, "_warned_cudnn", False):
            import warnings
            warnings.warn(f"CrossAttentionFusion.temporal_model: shape (batch {B}, hidden {lstm.hidden_size}, input {In}) is "
                          "outside the sm_100a BiLSTM kernels (hidden 256/512, batch <= 64/128); running torch.nn.LSTM (cuDNN)")
            self._warned_cudnn = True
        fused_seq, _ = lstm(fused)
        return fused_seq
