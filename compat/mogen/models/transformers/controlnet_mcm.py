"""mogen/models/transformers/controlnet_mcm.py:29-403 (imported by path in tools/m2d_test.py:19, tools/s2g_test.py)."""
from motioncraft_b200.condition_encoder import WavEncoder  # noqa: F401
from motioncraft_b200.modules import ControlT2MBlock, ControlT2MHalf_MCM  # noqa: F401
