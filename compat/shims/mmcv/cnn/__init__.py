"""mmcv.cnn: the parent MODELS registry of mogen/models/builder.py:1, build_norm_layer / build_activation_layer
(mogen/models/gnns/stgcn.py)."""
from torch import nn

from ..utils import Registry, build_from_cfg


def build_model_from_cfg(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


MODELS = Registry("model", build_func=build_model_from_cfg)

_NORMS = {"BN": ("bn", nn.BatchNorm2d), "BN1d": ("bn", nn.BatchNorm1d), "BN2d": ("bn", nn.BatchNorm2d),
          "BN3d": ("bn", nn.BatchNorm3d), "SyncBN": ("bn", nn.SyncBatchNorm), "GN": ("gn", nn.GroupNorm),
          "LN": ("ln", nn.LayerNorm), "IN": ("in", nn.InstanceNorm2d), "IN1d": ("in", nn.InstanceNorm1d)}


def build_norm_layer(cfg, num_features, postfix=""):
    cfg_ = dict(cfg)
    layer_type = cfg_.pop("type")
    if layer_type not in _NORMS:
        raise KeyError(f"Unrecognized norm type {layer_type}")
    abbr, cls = _NORMS[layer_type]
    requires_grad = cfg_.pop("requires_grad", True)
    cfg_.setdefault("eps", 1e-5)
    if layer_type == "GN":
        layer = cls(num_channels=num_features, **cfg_)
    else:
        layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


_ACTS = {"ReLU": nn.ReLU, "LeakyReLU": nn.LeakyReLU, "PReLU": nn.PReLU, "ReLU6": nn.ReLU6, "ELU": nn.ELU,
         "Sigmoid": nn.Sigmoid, "Tanh": nn.Tanh, "GELU": nn.GELU, "SiLU": nn.SiLU}


def build_activation_layer(cfg):
    cfg_ = dict(cfg)
    return _ACTS[cfg_.pop("type")](**cfg_)
