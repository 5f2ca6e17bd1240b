"""mogen/models/utils/gaussian_diffusion.py -- schedule tables and the re-hosted sampling loops."""
from motioncraft_b200.diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType, SpacedDiffusion,  # noqa: F401
                                        build_diffusion, get_named_beta_schedule, space_timesteps)
