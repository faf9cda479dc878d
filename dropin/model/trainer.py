"""Shim for `from model.trainer import MultimodalTrainer` (reference main.py:11)."""
from multimodal_av_model_b200.trainer import MultimodalTrainer  # noqa: F401
