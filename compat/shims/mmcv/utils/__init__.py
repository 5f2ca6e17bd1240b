"""mmcv.utils: Registry / build_from_cfg (mogen/datasets/builder.py:10, pipelines/compose.py, core/*/builder.py),
get_logger (mogen/utils/logger.py), collect_env / get_git_hash (mogen/utils/collect_env.py)."""
import inspect
import logging
import subprocess
import sys

from ..misc import is_list_of, is_seq_of, is_str, is_tuple_of, mkdir_or_exist  # noqa: F401


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
    if "type" not in cfg and (default_args is None or "type" not in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    elif inspect.isclass(obj_type) or inspect.isfunction(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
    try:
        return obj_cls(**args)
    except Exception as e:
        raise type(e)(f"{obj_cls.__name__}: {e}") from e


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self._name = name
        self._module_dict = {}
        self._children = {}
        self._scope = scope
        self.parent = parent
        if build_func is None:
            build_func = parent.build_func if parent is not None else build_from_cfg
        self.build_func = build_func
        if parent is not None:
            parent._children[name] = self

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def __repr__(self):
        return f"{type(self).__name__}(name={self._name}, items={sorted(self._module_dict)})"

    @property
    def name(self):
        return self._name

    @property
    def scope(self):
        return self._scope

    @property
    def module_dict(self):
        return self._module_dict

    @property
    def children(self):
        return self._children

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register_module(self, module_class, module_name=None, force=False):
        if not inspect.isclass(module_class) and not inspect.isfunction(module_class):
            raise TypeError(f"module must be a class or a function, but got {type(module_class)}")
        names = [module_class.__name__] if module_name is None else ([module_name] if isinstance(module_name, str) else module_name)
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self.name}")
            self._module_dict[n] = module_class

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if module is not None:
            self._register_module(module, name, force)
            return module

        def _register(cls):
            self._register_module(cls, name, force)
            return cls
        return _register


_loggers = {}


def get_logger(name, log_file=None, log_level=logging.INFO, file_mode="w"):
    logger = logging.getLogger(name)
    if name in _loggers:
        return logger
    handlers = [logging.StreamHandler()]
    if log_file is not None:
        handlers.append(logging.FileHandler(log_file, file_mode))
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    for h in handlers:
        h.setFormatter(fmt)
        h.setLevel(log_level)
        logger.addHandler(h)
    logger.setLevel(log_level)
    logger.propagate = False
    _loggers[name] = True
    return logger


def print_log(msg, logger=None, level=logging.INFO):
    if logger is None:
        print(msg)
    elif isinstance(logger, logging.Logger):
        logger.log(level, msg)
    elif logger != "silent":
        get_logger(logger).log(level, msg)


def collect_env():
    import torch
    env = {"sys.platform": sys.platform, "Python": sys.version.replace("\n", ""), "PyTorch": torch.__version__,
           "CUDA available": torch.cuda.is_available()}
    if torch.cuda.is_available():
        env["GPU 0"] = torch.cuda.get_device_name(0)
    from .. import __version__ as v
    env["MMCV"] = f"{v} (motioncraft_b200 stand-in)"
    return env


def get_git_hash(fallback="unknown", digits=None):
    try:
        sha = subprocess.check_output(["git", "rev-parse", "HEAD"], stderr=subprocess.DEVNULL).decode().strip()
    except (OSError, subprocess.CalledProcessError):
        sha = fallback
    return sha[:digits] if digits else sha
