"""mogen/models/transformers/stmogen.py -- the pieces of the STMoGen family available so far (motioncraft_b200/pathb.py)."""
from motioncraft_b200.pathb import (PoseDecoder, PoseEncoder, STMoGenTransformer, get_smplx_slice)  # noqa: F401
