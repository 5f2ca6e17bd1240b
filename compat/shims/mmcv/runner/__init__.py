"""mmcv.runner: BaseModule, load_checkpoint, get_dist_info, init_dist, wrap_fp16_model (tools/test.py:11-12, 99-103) and
import-only stubs for the training names mogen/apis/train.py, mogen/core/* pull in."""
import os
import re
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch import nn


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = init_cfg

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):
        self._is_init = True


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(launcher, backend="nccl", **kwargs):
    if launcher != "pytorch":
        raise NotImplementedError(f"launcher {launcher!r}: only 'pytorch' (torchrun, one process per GPU) is provided")
    rank = int(os.environ["RANK"])
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank % max(1, torch.cuda.device_count()))))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, **kwargs)


def _load_file(filename, map_location=None):
    if not os.path.isfile(filename):
        raise FileNotFoundError(f"{filename} can not be found.")
    try:
        return torch.load(filename, map_location=map_location, weights_only=False)
    except TypeError:                      # very old torch
        return torch.load(filename, map_location=map_location)


def load_state_dict(module, state_dict, strict=False, logger=None):
    res = module.load_state_dict(state_dict, strict=strict)
    missing = [k for k in getattr(res, "missing_keys", []) if "num_batches_tracked" not in k]
    unexpected = list(getattr(res, "unexpected_keys", []))
    msgs = []
    if unexpected:
        msgs.append("unexpected key in source state_dict: " + ", ".join(unexpected))
    if missing:
        msgs.append("missing keys in source state_dict: " + ", ".join(missing))
    if msgs:
        text = "The model and loaded state dict do not match exactly\n" + "\n".join(msgs)
        if strict:
            raise RuntimeError(text)
        (logger.warning if logger is not None else print)(text)
    return res


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None, revise_keys=((r"^module\.", ""),)):
    """Loads `checkpoint['state_dict']` (or the bare dict) into `model`, stripping the DataParallel `module.` prefix;
    returns the checkpoint dict.  Checkpoints of MotionDiffusion carry `model.`-prefixed denoiser keys, which match
    the architecture's `.model` attribute directly (tools/test.py:102-103)."""
    checkpoint = _load_file(filename, map_location)
    if not isinstance(checkpoint, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    state_dict = checkpoint.get("state_dict", checkpoint)
    metadata = getattr(state_dict, "_metadata", OrderedDict())
    for pat, rep in revise_keys:
        state_dict = OrderedDict((re.sub(pat, rep, k), v) for k, v in state_dict.items())
    state_dict._metadata = metadata
    from ..parallel import is_module_wrapper
    target = model.module if is_module_wrapper(model) else model
    load_state_dict(target, state_dict, strict, logger)
    return checkpoint


def wrap_fp16_model(model):
    raise NotImplementedError("fp16 wrapping is a training-side mmcv feature; the B200 path picks its operand formats "
                              "inside the CUDA library")


def _training_only(name):
    def _raise(*a, **k):
        raise NotImplementedError(f"mmcv.runner.{name} is training-side and not provided by this stand-in")

    class _Stub:
        def __init__(self, *a, **k):
            _raise()
    _Stub.__name__ = name
    return _Stub


build_optimizer = _training_only("build_optimizer")
build_runner = _training_only("build_runner")
OptimizerHook = _training_only("OptimizerHook")
Fp16OptimizerHook = _training_only("Fp16OptimizerHook")
DistSamplerSeedHook = _training_only("DistSamplerSeedHook")
EvalHook = _training_only("EvalHook")
DistEvalHook = _training_only("DistEvalHook")
HOOKS = None
