"""Shim for `from model.encoder import VisualEncoder, AudioEncoder` (reference main.py:8)."""
from multimodal_av_model_b200.encoders import AudioEncoder, VisualEncoder  # noqa: F401
