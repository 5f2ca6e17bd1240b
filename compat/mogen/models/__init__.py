"""mogen.models -- the B200 implementation under the reference's import path (mogen/models/__init__.py:1-7).

Model families outside the rewritten hot path that the unchanged tools still build through the registry (the evaluator
`T2MContrastiveModel_SMPLX` of tools/m2d_test.py:398, mogen/models/rnns) are imported from the reference checkout when
one is present; they register themselves in the registry below through `from ..builder import SUBMODULES`."""
import os

from .. import reference_root as _reference_root

_ref = _reference_root()
if _ref is not None:
    __path__.append(os.path.join(_ref, "mogen", "models"))      # rnns / gnns: the reference's files, found after ours

from .architectures import *  # noqa: F401,F403,E402
from .attentions import *  # noqa: F401,F403,E402
from .builder import *  # noqa: F401,F403,E402
from .losses import *  # noqa: F401,F403,E402
from .transformers import *  # noqa: F401,F403,E402
from .utils import *  # noqa: F401,F403,E402

try:                                                     # evaluator models (reference files, optional dependencies)
    from . import rnns  # noqa: F401,E402
except Exception:                                        # noqa: BLE001 - absent checkout or missing optional dependency
    rnns = None
