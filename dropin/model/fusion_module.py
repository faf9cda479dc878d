"""Shim for `from model.fusion_module import CrossAttentionFusion` (reference main.py:9)."""
from multimodal_av_model_b200.fusion_module import CrossAttentionFusion  # noqa: F401
