"""mogen/models/transformers/controlnet.py:340-423 -- the ControlNet wrapper of the STMoGen family (Path B).

tools/m2d_test.py:18 and tools/s2g_test.py import this name unconditionally but only construct it for
`cfg.model.model.type == 'STMoGenTransformer'` (m2d_test.py:374-377); configs/mcm/* take the ControlT2MHalf_MCM branch."""
from motioncraft_b200._lib import McmError


class ControlT2MHalf:
    def __init__(self, *args, **kwargs):
        raise McmError("ControlT2MHalf wraps STMoGenTransformer (configs/stmogen/*, SURVEY.md section 8 row f-1); the "
                       "configs/mcm/* models use ControlT2MHalf_MCM")
