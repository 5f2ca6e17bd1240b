"""mogen/models/builder.py:1-36 -- one registry under five names, `build_*` helpers; `None` config -> `None`."""
from motioncraft_b200.registry import (ARCHITECTURES, ATTENTIONS, LOSSES, MODELS, SUBMODULES, build_architecture,  # noqa: F401
                                       build_attention, build_loss, build_submodule)


def build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    return registry.build(cfg, default_args)


__all__ = ["MODELS", "LOSSES", "ARCHITECTURES", "SUBMODULES", "ATTENTIONS", "build_loss", "build_architecture",
           "build_submodule", "build_attention", "build_from_cfg"]
