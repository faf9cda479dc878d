"""GPU: batched beam-search kernel, bit-exact token lists against the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_cases

pytestmark = pytest.mark.gpu
BEAM = {k: v for k, v in load_cases("beam_cases.npz").items() if not k.startswith("topk")}


def _pkg():
    import multimodal_av_model_b200 as pkg
    return pkg


@pytest.mark.parametrize("name", sorted(BEAM))
@pytest.mark.parametrize("fast", [1, 0])
def test_beam_matches_reference_fixture(name, fast):
    pkg = _pkg()
    c = BEAM[name]
    pkg._lib.set_tuning("beam_fast", fast)
    try:
        ids = pkg.simple_beam_search(torch.from_numpy(c["lp"]).cuda(), beam_width=int(c["beam"]), blank=int(c["blank"]))
    finally:
        pkg._lib.set_tuning("beam_fast", 1)
    assert ids == c["ids"].tolist()


@pytest.mark.parametrize("V,k", [(800, 16), (800, 13), (100, 5), (64, 32), (33, 32)])
def test_beam_wide_beam_tie_route(V, k):
    """beam*64 > V: torch.topk switches to nth_element + sort (TopKImpl.h:45,66-76); tied rows must still
    come out in its order (checked through the debug export of ALL beams against the oracle)."""
    pkg = _pkg()
    g = torch.Generator().manual_seed(V + k)
    lp = (0.2 * torch.randn(3, 24, V, generator=g)).log_softmax(-1).bfloat16().float()
    lp[1, 5] = -3.0
    res, scores, paths = pkg.beam_search_batch(lp.cuda(), beam_width=k, blank=3, return_debug=True)
    for i in range(3):
        ids, sc, pa = oracle.beam_search(lp[i].numpy(), k, 3, debug=True)
        assert res[i] == ids
        assert np.array_equal(paths[i].numpy(), pa)
        assert np.array_equal(scores[i].numpy(), sc)


def test_beam_batch_vs_oracle_with_debug_export():
    pkg = _pkg()
    g = torch.Generator().manual_seed(0)
    N, T, V, k = 37, 150, 800, 10
    lp = (3 * torch.randn(N, T, V, generator=g)).log_softmax(-1)
    # sprinkle exact ties into a third of the utterances
    lp[::3] = lp[::3].bfloat16().float()
    res, scores, paths = pkg.beam_search_batch(lp.cuda(), beam_width=k, blank=3, return_debug=True)
    for i in range(N):
        ids, sc, pa = oracle.beam_search(lp[i].numpy(), k, 3, debug=True)
        assert res[i] == ids
        assert np.array_equal(scores[i].numpy(), sc)          # fp64 sums, same order of additions
        assert np.array_equal(paths[i].numpy(), pa)
    # default (no debug) path returns the same lists
    assert pkg.beam_search_batch(lp.cuda(), beam_width=k, blank=3) == res


@pytest.mark.parametrize("V,k", [(1024, 10), (1000, 10), (993, 5), (500, 10), (832, 11), (31, 5)])
def test_beam_vocab_shapes_cover_every_topk_mode(V, k):
    """Top-k kernel template modes: V = 32*NV exactly (1024, 832), one ragged slot (1000, 993), generic (500, 31),
    all with exact ties (bf16-rounded values, one constant row, one row with -inf), all beams vs the oracle."""
    pkg = _pkg()
    g = torch.Generator().manual_seed(V * 31 + k)
    lp = (2 * torch.randn(4, 40, V, generator=g)).log_softmax(-1)
    lp[1] = lp[1].bfloat16().float()
    lp[2, 7] = -2.5
    lp[3, 3, : V // 2] = float("-inf")
    res, scores, paths = pkg.beam_search_batch(lp.cuda(), beam_width=k, blank=0, return_debug=True)
    for i in range(4):
        ids, sc, pa = oracle.beam_search(lp[i].numpy(), k, 0, debug=True)
        assert res[i] == ids
        assert np.array_equal(paths[i].numpy(), pa)
        assert np.array_equal(scores[i].numpy(), sc)


def test_beam_lengths_long_T_and_strides():
    pkg = _pkg()
    g = torch.Generator().manual_seed(1)
    N, T, V = 5, 700, 801                      # back-pointers spill to the global workspace (T*beam*4 > 24 KiB)
    lp = (2 * torch.randn(N, T, V + 3, generator=g)).log_softmax(-1)[:, :, :V]   # row stride != V, misaligned rows
    lens = torch.tensor([700, 1, 0, 350, 699])
    res = pkg.beam_search_batch(lp.cuda(), beam_width=10, blank=0, lengths=lens)
    for i in range(N):
        assert res[i] == oracle.beam_search(lp[i, :int(lens[i])].contiguous().numpy(), 10, 0)
    assert res[2] == []


def test_beam_host_input_streams_in_chunks():
    """Host-resident log-probs are decoded on the GPU through a chunked copy/decode pipeline: same token lists as
    the device-resident call (pinned and pageable sources, ragged last chunk, per-utterance lengths, debug export)."""
    pkg = _pkg()
    from multimodal_av_model_b200.beam_search import _beam_search_from_host
    g = torch.Generator().manual_seed(3)
    N, T, V = 23, 30, 800
    lp = (3 * torch.randn(N, T, V, generator=g)).log_softmax(-1)
    lens = torch.randint(0, T + 1, (N,), generator=g)
    want = pkg.beam_search_batch(lp.cuda(), beam_width=10, blank=3)
    want_l = pkg.beam_search_batch(lp.cuda(), beam_width=10, blank=3, lengths=lens)
    for src in (lp, lp.pin_memory()):
        assert pkg.beam_search_batch(src, beam_width=10, blank=3) == want                   # one chunk
        got = _beam_search_from_host(src, 10, 3, None, False, chunk_bytes=5 * T * V * 4)    # 5 utterances per chunk
        assert got == want
        assert _beam_search_from_host(src, 10, 3, lens, False, chunk_bytes=4 * T * V * 4) == want_l
    r, sc, pa = _beam_search_from_host(lp, 10, 3, None, True, chunk_bytes=6 * T * V * 4)
    r0, sc0, pa0 = pkg.beam_search_batch(lp.cuda(), beam_width=10, blank=3, return_debug=True)
    assert r == r0 and torch.equal(sc, sc0) and torch.equal(pa, pa0)
    assert pkg.simple_beam_search(lp[0], beam_width=5, blank=3) == pkg.simple_beam_search(lp[0].cuda(), beam_width=5, blank=3)


def test_fast_decode_and_errors():
    pkg = _pkg()

    class Tok:
        id_to_token = ["<unk>", "<s>", "</s>", "<blank>", "▁", "a", "b"]
        vocab_size = 7
        blank_id = 3
    assert pkg.fast_decode([5, 4, 6, 3, 99, -1, 4], Tok()) == "a b"
    with pytest.raises(RuntimeError):
        pkg.simple_beam_search(torch.zeros(4, 6).cuda(), beam_width=7, blank=0)


def test_beam_config5_full_size_equals_collapsed_row_maximum():
    """BASELINE config 5 at full size (4096 x [150, 800], beam 10).  Size-independent property of the reference's
    search (beam_search.py:13-40): candidate (beam 0, k 0) is always inserted first and fp addition
    is monotone, so the winning path is the per-frame first top-k index, i.e. the row argmax wherever the maximum is
    unique; the returned ids are its CTC collapse.  Checked for all 4096 utterances against torch ops on the GPU, and
    for a sample of them against the oracle."""
    pkg = _pkg()
    N, T, V, k, blank = 4096, 150, 800, 10, 3
    g = torch.Generator(device="cuda").manual_seed(7)
    lp = (3 * torch.randn(N, T, V, generator=g, device="cuda")).log_softmax(-1)
    res = pkg.beam_search_batch(lp, beam_width=k, blank=blank)
    top2 = lp.topk(2, dim=-1).values
    unique_max = (top2[..., 0] > top2[..., 1]).all(dim=1).cpu().numpy()                   # [N]
    assert unique_max.mean() > 0.9
    am = lp.argmax(-1).cpu().numpy()                                                      # [N,T]
    checked = 0
    for i in range(N):
        if not unique_max[i]:
            continue
        row = am[i]
        keep = (row != blank) & np.concatenate(([True], row[1:] != row[:-1]))
        assert res[i] == row[keep].tolist(), i
        checked += 1
    assert checked > 3600
    for i in range(0, N, 512):
        assert res[i] == oracle.beam_search(lp[i].cpu().numpy(), k, blank)
    # decoding is a pure function of the utterance: any sub-batch, in any order, gives the same lists
    sub = torch.tensor([4095, 17, 2048, 17], device="cuda")
    assert pkg.beam_search_batch(lp[sub], beam_width=k, blank=blank) == [res[4095], res[17], res[2048], res[17]]


@pytest.fixture
def fused_knobs():
    pkg = _pkg()
    yield pkg._lib.set_tuning
    for key, dflt in (("beam_fused", -1), ("beam_fused_grid", 0)):
        pkg._lib.set_tuning(key, dflt)


@pytest.mark.parametrize("grid", [0, 3, 1])
def test_beam_fused_kernel_all_beams_vs_oracle(grid, fused_knobs):
    """Fused decode kernel (top-k warps + recurrence warps in one persistent CTA, lists in shared memory): all beams,
    scores and paths against the oracle, exact ties included; one utterance per CTA, and 13 / 37 utterances through the
    two list buffers of a CTA."""
    pkg = _pkg()
    fused_knobs("beam_fused", 1); fused_knobs("beam_fused_grid", grid)
    g = torch.Generator().manual_seed(10 + grid)
    N, T, V, k = 37, 150, 800, 10
    lp = (3 * torch.randn(N, T, V, generator=g)).log_softmax(-1)
    lp[::3] = lp[::3].bfloat16().float()
    lp[4, 9] = -2.0                                       # a constant row: torch.topk's tie order
    res, scores, paths = pkg.beam_search_batch(lp.cuda(), beam_width=k, blank=3, return_debug=True)
    for i in range(N):
        ids, sc, pa = oracle.beam_search(lp[i].numpy(), k, 3, debug=True)
        assert res[i] == ids
        assert np.array_equal(scores[i].numpy(), sc)
        assert np.array_equal(paths[i].numpy(), pa)


@pytest.mark.parametrize("V,k", [(801, 10), (832, 11), (790, 10), (600, 7), (513, 1)])
def test_beam_fused_kernel_vocab_modes_lengths_strides(V, k, fused_knobs):
    """Every template mode of the fused kernel (V = 32*NV, one ragged slot, generic), ragged lengths (0, 1, T), rows with
    a stride != V, few CTAs so that list buffers are reused many times: identical to the two-phase kernels, and to the
    oracle on a sample."""
    pkg = _pkg()
    g = torch.Generator().manual_seed(V * 7 + k)
    N, T = 61, 40
    lp = (2 * torch.randn(N, T, V + 5, generator=g)).log_softmax(-1)[:, :, :V]
    lp[1::4] = lp[1::4].bfloat16().float()
    lp[2, 7] = -2.5
    lp[3, 3, : V // 2] = float("-inf")
    lens = torch.randint(0, T + 1, (N,), generator=g)
    lens[0], lens[5], lens[6] = T, 0, 1
    dev = lp.cuda()
    fused_knobs("beam_fused", 0)
    want = pkg.beam_search_batch(dev, beam_width=k, blank=0, lengths=lens)
    want_full = pkg.beam_search_batch(dev, beam_width=k, blank=0)
    fused_knobs("beam_fused", 1)
    for grid in (2, 5, 0):
        fused_knobs("beam_fused_grid", grid)
        assert pkg.beam_search_batch(dev, beam_width=k, blank=0, lengths=lens) == want
        assert pkg.beam_search_batch(dev, beam_width=k, blank=0) == want_full
    for i in range(0, N, 7):
        assert want[i] == oracle.beam_search(lp[i, :int(lens[i])].contiguous().numpy(), k, 0)
    assert want[5] == []


def test_beam_auto_policy_is_value_neutral(fused_knobs):
    """The default route (fused kernel up to 16 x SMs utterances, two-phase kernels above) returns what either forced
    route returns."""
    pkg = _pkg()
    g = torch.Generator(device="cuda").manual_seed(11)
    lp = (3 * torch.randn(700, 60, 800, generator=g, device="cuda")).log_softmax(-1)
    assert pkg._lib.lib().avctc_beam_route(700, 60, 800, 10) == 3 and pkg._lib.lib().avctc_beam_route(4096, 150, 800, 10) == 2
    auto = pkg.beam_search_batch(lp, beam_width=10, blank=3)
    for forced in (0, 1):
        fused_knobs("beam_fused", forced)
        assert pkg.beam_search_batch(lp, beam_width=10, blank=3) == auto
