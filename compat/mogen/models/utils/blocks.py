"""mogen/models/utils/blocks.py:11-71 (WavEncoder of the speech ControlNet branch)."""
from motioncraft_b200.condition_encoder import BasicBlock, WavEncoder  # noqa: F401
