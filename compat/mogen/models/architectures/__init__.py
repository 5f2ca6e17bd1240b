from .base_architecture import BaseArchitecture  # noqa: F401
from .diffusion_architecture import MotionDiffusion  # noqa: F401

__all__ = ["BaseArchitecture", "MotionDiffusion"]
