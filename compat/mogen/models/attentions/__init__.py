from .efficient_attention import EfficientCrossAttention, EfficientSelfAttention  # noqa: F401

__all__ = ["EfficientSelfAttention", "EfficientCrossAttention"]
