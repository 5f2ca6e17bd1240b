"""mogen/models/utils/stylization_block.py:14-40."""
from motioncraft_b200.modules import StylizationBlock  # noqa: F401
