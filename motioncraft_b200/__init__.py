"""motioncraft_b200 -- B200-native (sm_100a) denoising hot path of cure-lab/MotionCraft (configs/mcm/*).

Public surface (mirrors `mogen.models`): build_architecture / build_submodule / build_attention /
build_loss and the registered types MotionDiffusion, MCMTransformer, ControlT2MHalf_MCM,
EfficientSelfAttention, EfficientCrossAttention, MSELoss.
"""
__version__ = "0.1.0"

from .registry import (ARCHITECTURES, ATTENTIONS, LOSSES, MODELS, SUBMODULES, build_architecture,  # noqa: F401
                       build_attention, build_loss, build_submodule)
from . import modules as _modules  # noqa: F401,E402  (registers the model types)
from . import architecture as _architecture  # noqa: F401,E402
from .architecture import MotionDiffusion  # noqa: F401,E402
from .modules import (ControlT2MHalf_MCM, DecoderLayer, EfficientCrossAttention,  # noqa: F401,E402
                      EfficientSelfAttention, MCMTransformer)
from .diffusion import GaussianDiffusion, SpacedDiffusion, space_timesteps  # noqa: F401,E402
