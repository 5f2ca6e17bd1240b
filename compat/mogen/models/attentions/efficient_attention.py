"""mogen/models/attentions/efficient_attention.py:9-92."""
from motioncraft_b200.modules import EfficientCrossAttention, EfficientSelfAttention  # noqa: F401
