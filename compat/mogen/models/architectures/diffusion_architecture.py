"""mogen/models/architectures/diffusion_architecture.py:25-204."""
from motioncraft_b200.architecture import MotionDiffusion  # noqa: F401
from motioncraft_b200.diffusion import build_diffusion  # noqa: F401
