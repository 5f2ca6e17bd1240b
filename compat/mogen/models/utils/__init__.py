"""mogen.models.utils -- rewritten pieces come from motioncraft_b200 (gaussian_diffusion, stylization_block, scheduler,
blocks); helper modules the kept-as-is subsystems import from here (quaternion, word_vectorizer, misc, ...) are the
reference's own files, found through the search path below when a checkout is present."""
import os

from ... import reference_root as _reference_root

_ref = _reference_root()
if _ref is not None:
    __path__.append(os.path.join(_ref, "mogen", "models", "utils"))

from .gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType, SpacedDiffusion,  # noqa: F401,E402
                                 get_named_beta_schedule, space_timesteps)
from .stylization_block import StylizationBlock  # noqa: F401,E402

__all__ = ["GaussianDiffusion", "SpacedDiffusion", "ModelMeanType", "ModelVarType", "LossType", "space_timesteps",
           "get_named_beta_schedule", "StylizationBlock"]
