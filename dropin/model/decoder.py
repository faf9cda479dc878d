"""Shim for `from model.decoder import CTCDecoder` (reference main.py:10)."""
from multimodal_av_model_b200.decoder import CTCDecoder  # noqa: F401
