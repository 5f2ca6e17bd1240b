"""mogen/models/transformers/diffusion_transformer.py:15-238.  configs/mcm/* instantiate the MCMTransformer subclass
only; the base-class name is kept for isinstance checks and imports."""
from motioncraft_b200.modules import FFN, MCMTransformer  # noqa: F401

DiffusionTransformer = MCMTransformer
