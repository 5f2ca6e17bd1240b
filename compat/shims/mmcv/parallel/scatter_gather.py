"""scatter / scatter_kwargs with DataContainer support (mmcv.parallel.scatter_gather), one target device."""
import torch


def _to(obj, device):
    from . import DataContainer
    if isinstance(obj, torch.Tensor):
        return obj if device is None else obj.to(device, non_blocking=True)
    if isinstance(obj, DataContainer):
        # collate() produced one entry per GPU: take ours.  cpu_only data stays on the host, unwrapped.
        data = obj.data[0] if _is_collated(obj) and len(obj.data) > 0 else obj.data
        return data if obj.cpu_only else _to(data, device)
    if isinstance(obj, tuple) and len(obj) > 0:
        return tuple(_to(o, device) for o in obj)
    if isinstance(obj, list) and len(obj) > 0:
        return [_to(o, device) for o in obj]
    if isinstance(obj, dict) and len(obj) > 0:
        return type(obj)((k, _to(v, device)) for k, v in obj.items())
    return obj


def _is_collated(dc):
    # after collate() the payload is a list with one element per GPU (a tensor, or a list of per-sample objects)
    return isinstance(dc.data, list)


def scatter(inputs, target_gpus, dim=0):
    if len(target_gpus) != 1:
        raise NotImplementedError("one process drives one device")
    dev = None if target_gpus[0] == -1 else torch.device("cuda", target_gpus[0])
    return [_to(inputs, dev)]


def scatter_kwargs(inputs, kwargs, target_gpus, dim=0):
    inputs = scatter(inputs, target_gpus, dim) if inputs else []
    kwargs = scatter(kwargs, target_gpus, dim) if kwargs else []
    if len(inputs) < len(kwargs):
        inputs.extend([() for _ in range(len(kwargs) - len(inputs))])
    elif len(kwargs) < len(inputs):
        kwargs.extend([{} for _ in range(len(inputs) - len(kwargs))])
    return tuple(inputs), tuple(kwargs)
