from .controlnet import ControlT2MHalf  # noqa: F401
from .controlnet_mcm import ControlT2MBlock, ControlT2MHalf_MCM  # noqa: F401
from .diffusion_transformer import FFN, DiffusionTransformer  # noqa: F401
from .mcm import DecoderLayer, MCMTransformer  # noqa: F401

__all__ = ["MCMTransformer", "DecoderLayer", "ControlT2MHalf_MCM", "ControlT2MBlock", "ControlT2MHalf",
           "DiffusionTransformer", "FFN"]
