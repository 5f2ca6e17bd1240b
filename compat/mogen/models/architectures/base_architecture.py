"""mogen/models/architectures/base_architecture.py -- the eval-path piece (`split_results`, :112-140) lives on
`MotionDiffusion` in the B200 package; the base-class name is kept for imports and isinstance checks."""
from torch import nn

from motioncraft_b200.architecture import MotionDiffusion as _MotionDiffusion


class BaseArchitecture(nn.Module):
    split_results = staticmethod(_MotionDiffusion.split_results)
