"""Shim for `from beam_search import simple_beam_search, fast_decode` (reference model/trainer.py:8)."""
from multimodal_av_model_b200.beam_search import beam_search_batch, fast_decode, simple_beam_search  # noqa: F401
