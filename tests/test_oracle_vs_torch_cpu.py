"""CPU: the oracle pinned LIVE against the arithmetic the reference calls -- torch's own CPU kernels in this image
(nn.CTCLoss at /root/reference/model/trainer.py:25,116-117; torch.topk at /root/reference/beam_search.py:15) -- on
seeded random cases, next to the committed fixtures of tests/test_oracle_golden.py.  Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

import oracle


def _case(rng, T, B, V, blank, lmax, peaked):
    z = rng.standard_normal((T, B, V)) * (8.0 if peaked else 1.0)
    lp = z - np.log(np.exp(z - z.max(-1, keepdims=True)).sum(-1, keepdims=True)) - z.max(-1, keepdims=True)
    tl = rng.integers(0, lmax + 1, size=B)
    il = rng.integers(0, T + 1, size=B)
    il[0] = T
    ids = np.array([c for c in range(V) if c != blank])
    tg = np.zeros((B, max(lmax, 1)), dtype=np.int64)
    for b in range(B):
        row = rng.choice(ids, size=tl[b])
        for j in range(1, tl[b]):
            if rng.random() < 0.3:
                row[j] = row[j - 1]                  # repeats: need a blank in between, many samples infeasible
        tg[b, :tl[b]] = row
    return lp, tg, il.astype(np.int64), tl.astype(np.int64)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_ctc_oracle_equals_torch_float64(seed, reduction):
    """Loss, per-sample nll and the softmax-folded gradient, with ragged input lengths (0 included), empty targets,
    repeated labels and infeasible samples, zero_infinity on and off."""
    rng = np.random.default_rng(100 + seed)
    T, B = int(rng.integers(1, 40)), int(rng.integers(1, 6))
    V, blank = (int(rng.integers(3, 30)), 0) if seed % 2 else (int(rng.integers(5, 30)), 3)
    lp, tg, il, tl = _case(rng, T, B, V, blank, int(rng.integers(0, 12)), peaked=seed % 3 == 0)
    zi = bool(seed % 2)
    x = torch.from_numpy(lp).double().requires_grad_()
    ref = torch.nn.functional.ctc_loss(x, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl), blank=blank,
                                       reduction=reduction, zero_infinity=zi)
    w = torch.linspace(0.5, 1.5, ref.numel(), dtype=torch.float64).reshape(ref.shape)
    got = oracle.ctc_loss(lp, tg, il, tl, blank=blank, reduction=reduction, zero_infinity=zi)
    mine = np.asarray(got["nll"] if reduction == "none" else got["loss"], dtype=np.float64)
    if reduction == "none" and zi:
        mine = np.where(np.isinf(mine), 0.0, mine)          # torch zeroes the infinite entries it returns
    r = ref.detach().numpy()
    assert np.array_equal(np.isfinite(mine), np.isfinite(r))
    fin = np.isfinite(r)
    assert np.allclose(mine[fin], r[fin], rtol=1e-11, atol=1e-11)
    if not np.all(fin) and not zi:
        return                                               # torch's gradient of an infinite loss is NaN/garbage
    if reduction == "none":
        (ref * w).sum().backward()
        # the oracle returns the gradient of sum_b nll_b for 'none'; weight it per sample like the reference side
        g = got["grad"] * w.numpy()[None, :, None]
        want = x.grad.numpy()
        assert np.abs(g - want).max() <= 1e-10 * max(np.abs(want).max(), 1e-30)
    else:
        ref.backward()
        want = x.grad.numpy()
        assert np.abs(got["grad"] - want).max() <= 1e-10 * max(np.abs(want).max(), 1e-30)


@pytest.mark.parametrize("V,k", [(800, 5), (800, 10), (801, 10), (800, 13), (801, 16), (64, 1), (64, 5), (33, 32), (12, 12)])
def test_topk_oracle_equals_torch_on_tie_heavy_rows(V, k):
    """Both algorithms of TopKImpl.h (partial_sort when k*64 <= V, nth_element + sort otherwise), rows quantised so
    that ties are everywhere, plus NaN / inf entries."""
    rng = np.random.default_rng(V * 100 + k)
    for trial in range(25):
        levels = int(rng.integers(1, 9))
        row = rng.integers(0, levels, size=V).astype(np.float32) * 0.5 - 3.0
        if trial % 5 == 1:
            row[rng.integers(0, V, size=3)] = np.inf
        if trial % 5 == 2:
            row[rng.integers(0, V, size=2)] = -np.inf
        if trial % 5 == 3:
            row[rng.integers(0, V, size=2)] = np.nan
        tv, ti = torch.topk(torch.from_numpy(row), k)
        vals, idx = oracle.topk(row, k)
        assert idx.tolist() == ti.tolist(), (trial, levels)
        assert np.array_equal(vals, tv.numpy(), equal_nan=True)


def test_beam_oracle_equals_collapsed_torch_topk_with_ties():
    """beam_search.py:13-40 reduces to the CTC collapse of torch.topk(row, beam).indices[0] per frame;
    here with bf16-quantised log-probs so that the row maximum is tied in many frames and torch.topk's order decides."""
    g = torch.Generator().manual_seed(3)
    for V, k, blank in ((800, 5, 3), (800, 10, 3), (801, 10, 0), (64, 16, 3)):
        lp = (2 * torch.randn(120, V, generator=g)).log_softmax(-1).bfloat16().float()
        lp[::7] = lp[::7].round()                            # frames with massive ties
        first = [int(torch.topk(lp[t], k).indices[0]) for t in range(lp.shape[0])]
        want, prev = [], None
        for c in first:
            if c != prev and c != blank:
                want.append(c)
            prev = c
        assert oracle.beam_search(lp.numpy(), k, blank) == want
