"""Run the CTC / fusion / beam parity tests with every fallback knob flipped (one at a time): the alternative kernels
and launch modes must stay parity-green too."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
import pytest
from multimodal_av_model_b200 import _lib
DEFAULTS = {"pdl": 1, "ctc_ws": 1, "ctc_lin": 1, "ctc_pf": 1, "ctc_overlap": 1, "ctc_stage": 1, "beam_two_phase": 1, "beam_pf": 1, "beam_fused": -1,
            "lstm_groups": 0}
os.environ["AVCTC_KNOB_MATRIX"] = "1"      # tests that assert a launch mode was taken skip that assertion
CASES = [("pdl", 0, ["tests/test_ctc_gpu.py", "tests/test_fusion_gpu.py", "tests/test_attention_gpu.py", "tests/test_ctc_head_gpu.py"]),
         ("ctc_ws", 0, ["tests/test_ctc_gpu.py"]),
         ("ctc_lin", 0, ["tests/test_ctc_gpu.py"]),
         ("ctc_pf", 0, ["tests/test_ctc_gpu.py"]),
         ("ctc_overlap", 0, ["tests/test_ctc_gpu.py"]),
         ("ctc_stage", 0, ["tests/test_ctc_gpu.py"]),
         ("lstm_groups", 1, ["tests/test_lstm_gpu.py", "tests/test_bench_sizes_gpu.py"]),
         ("beam_two_phase", 0, ["tests/test_beam_gpu.py"]),
         ("beam_pf", 0, ["tests/test_beam_gpu.py"]),
         ("beam_fused", 0, ["tests/test_beam_gpu.py"])]
bad = 0
for key, val, files in CASES:
    for k, v in DEFAULTS.items():
        _lib.set_tuning(k, v)
    _lib.set_tuning(key, val)
    rc = pytest.main(["-x", "-q", "-p", "no:cacheprovider"] + [os.path.join(root, f) for f in files])
    print(f"### {key}={val}: rc={int(rc)}", flush=True)
    bad += int(rc) != 0
for k, v in DEFAULTS.items():
    _lib.set_tuning(k, v)
print("### knob matrix:", "ok" if bad == 0 else f"{bad} failing configurations")
sys.exit(1 if bad else 0)
