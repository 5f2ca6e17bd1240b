"""`mogen` -- import-path compatibility shell of motioncraft_b200.

North star: "keeping the mogen model-registry / forward_test API so configs/mcm/* and tools/{test,visualize,m2d_*,s2g_*}.py
run unchanged"; subsystems rewritten: mogen/models/{attentions, transformers, architectures}, the diffusion scheduler;
`mogen/apis`, `mogen/datasets`, `mogen/core`, `mogen/utils` and `tools/*` stay as the reference has them.

This package therefore provides `mogen.models.*` from the B200 implementation and lets every OTHER `mogen.<sub>` import
fall through to an unmodified checkout of the reference: `MOGEN_REFERENCE_ROOT` (or the current working directory, which
is where the reference's tools are run from) is appended to this package's search path.  Nothing of the reference is
copied.  Mirrors mogen/__init__.py:1-56 (version export, `digit_version`, the mmcv version gate).
"""
import os
import warnings

import mmcv
from packaging.version import parse

__version__ = "0.0.1+b200"


def digit_version(version_str: str, length: int = 4):
    """Version string -> tuple of ints for comparisons (alpha < beta < rc), as mogen/__init__.py:9-42."""
    version = parse(version_str)
    assert version.release, f"failed to parse version {version_str}"
    release = (list(version.release)[:length] + [0] * length)[:length]
    if version.is_prerelease:
        mapping = {"a": -3, "b": -2, "rc": -1}
        if version.pre and version.pre[0] in mapping:
            release.extend([mapping[version.pre[0]], version.pre[-1]])
        else:
            if version.pre:
                warnings.warn(f"unknown prerelease version {version.pre[0]}, version checking may go wrong")
            release.extend([-4, 0])
    elif version.is_postrelease:
        release.extend([1, version.post])
    else:
        release.extend([0, 0])
    return tuple(release)


mmcv_minimum_version, mmcv_maximum_version = "1.4.2", "1.9.0"
mmcv_version = digit_version(mmcv.__version__)
assert digit_version(mmcv_minimum_version) <= mmcv_version <= digit_version(mmcv_maximum_version), \
    f"MMCV=={mmcv.__version__} is used but incompatible. Please install mmcv>={mmcv_minimum_version}, <={mmcv_maximum_version}."


def reference_root():
    """Directory of an unmodified reference checkout whose `mogen/{apis,datasets,core,utils}` are used as they are."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("MOGEN_REFERENCE_ROOT"), os.getcwd()):
        if cand and os.path.isdir(os.path.join(cand, "mogen", "apis")) and \
                os.path.abspath(os.path.join(cand, "mogen")) != here:
            return os.path.abspath(cand)
    return None


_ref = reference_root()
if _ref is not None:
    __path__.append(os.path.join(_ref, "mogen"))       # apis / datasets / core / utils: the reference's own files

__all__ = ["__version__", "digit_version"]
