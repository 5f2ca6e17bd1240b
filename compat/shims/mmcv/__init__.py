"""Minimal stand-in for the mmcv 1.x surface that the UNCHANGED callers of the MotionCraft hot path import
(tools/{test,visualize,m2d_*,s2g_*}.py, mogen/apis, mogen/datasets, mogen/core, mogen/utils).

mmcv-full 1.x is not installable in the target image (no wheel, no network).  This package is NOT a port of mmcv: it
implements, from mmcv's documented behaviour, exactly the calls those files make -- `Config.fromfile` with `_base_`
inheritance, `DictAction`, `ProgressBar`, `dump` / `load`, `MMDataParallel` with the `DataContainer(cpu_only=True)`
unwrap, `collate`, `load_checkpoint`, `Registry` / `build_from_cfg`, `BaseModule`, `get_dist_info`, `init_dist` -- and
stubs that raise for the training-only names those modules merely import.  Put it on PYTHONPATH only where the real
mmcv is absent; the real package, when present, is preferred.
"""
__version__ = "1.7.0"      # inside the range mogen/__init__.py:46-54 asserts (>= 1.4.2, <= 1.9.0)

from .config import Config, ConfigDict, DictAction  # noqa: F401
from .fileio import dump, load  # noqa: F401
from .misc import ProgressBar, is_list_of, is_seq_of, is_str, is_tuple_of, mkdir_or_exist  # noqa: F401
from . import cnn, parallel, runner, utils  # noqa: F401,E402
from .parallel import DataContainer  # noqa: F401,E402
