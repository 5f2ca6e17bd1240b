"""Model registry with the call surface of `mogen/models/builder.py:1-36`.

The reference aliases ONE mmcv registry as MODELS / LOSSES / ARCHITECTURES / SUBMODULES / ATTENTIONS
and builds objects from `dict(type=..., **kwargs)` configs; mmcv is not available here, so this is a
self-contained registry with the same names and build semantics (`None` config -> `None`).
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, default_args=None):
        if cfg is None:
            return None
        if not isinstance(cfg, dict) and hasattr(cfg, "items"):
            cfg = dict(cfg.items())
        if "type" not in cfg:
            raise KeyError(f"config for registry {self.name} needs a 'type' key, got {sorted(cfg)}")
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
LOSSES = MODELS
ARCHITECTURES = MODELS
SUBMODULES = MODELS
ATTENTIONS = MODELS


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_architecture(cfg):
    return ARCHITECTURES.build(cfg)


def build_submodule(cfg):
    return SUBMODULES.build(cfg)


def build_attention(cfg):
    return ATTENTIONS.build(cfg)
