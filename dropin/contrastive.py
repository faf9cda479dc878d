"""Shim for `from contrastive import contrastive_loss_with_mask` (reference model/trainer.py:7)."""
from multimodal_av_model_b200.contrastive import (TEMPERATURE, WEIGHT_NEG_SUPPRESS, WEIGHT_POS_ALIGN,  # noqa: F401
                                                   contrastive_loss_with_mask)
