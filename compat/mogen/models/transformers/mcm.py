"""mogen/models/transformers/mcm.py:12-102."""
from motioncraft_b200.modules import DecoderLayer, MCMTransformer  # noqa: F401
