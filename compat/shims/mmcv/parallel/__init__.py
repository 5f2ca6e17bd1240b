"""mmcv.parallel: DataContainer, collate, scatter_kwargs, MMDataParallel, MMDistributedDataParallel, MODULE_WRAPPERS.

The behaviour the unchanged callers rely on: `Collect` wraps `motion_metas` in
`DataContainer(cpu_only=True)` (mogen/datasets/pipelines/formatting.py:98); `collate` batches such containers into a
container holding a list (per GPU) of lists (per sample); `MMDataParallel.forward` scatters the keyword arguments to
device_ids[0] -- tensors are moved, a cpu_only container is UNWRAPPED to the plain per-sample list that
`MotionDiffusion.forward` indexes (diffusion_architecture.py:102-104)."""
from collections.abc import Mapping, Sequence

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils.data.dataloader import default_collate

from ..utils import Registry
from . import scatter_gather  # noqa: F401


class DataContainer:
    def __init__(self, data, stack=False, padding_value=0, cpu_only=False, pad_dims=2):
        self._data, self._cpu_only, self._stack, self._padding_value = data, cpu_only, stack, padding_value
        assert pad_dims in (None, 1, 2, 3)
        self._pad_dims = pad_dims

    def __repr__(self):
        return f"{type(self).__name__}({self.data!r})"

    def __len__(self):
        return len(self._data)

    @property
    def data(self):
        return self._data

    @property
    def datatype(self):
        return self.data.type() if isinstance(self.data, torch.Tensor) else type(self.data)

    @property
    def cpu_only(self):
        return self._cpu_only

    @property
    def stack(self):
        return self._stack

    @property
    def padding_value(self):
        return self._padding_value

    @property
    def pad_dims(self):
        return self._pad_dims

    def size(self, *a, **k):
        return self.data.size(*a, **k)

    def dim(self):
        return self.data.dim()


def collate(batch, samples_per_gpu=1):
    """Puts each data field into a tensor / DataContainer with outer dimension batch size (mmcv semantics)."""
    if not isinstance(batch, Sequence):
        raise TypeError(f"{type(batch)} is not supported.")
    if isinstance(batch[0], DataContainer):
        stacked = []
        if batch[0].cpu_only:
            for i in range(0, len(batch), samples_per_gpu):
                stacked.append([s.data for s in batch[i:i + samples_per_gpu]])
            return DataContainer(stacked, batch[0].stack, batch[0].padding_value, cpu_only=True)
        if batch[0].stack:
            for i in range(0, len(batch), samples_per_gpu):
                group = batch[i:i + samples_per_gpu]
                assert isinstance(group[0].data, torch.Tensor)
                if group[0].pad_dims is not None:
                    nd, pd = group[0].dim(), group[0].pad_dims
                    assert nd > pd
                    max_shape = [0] * pd
                    for d in range(1, pd + 1):
                        max_shape[d - 1] = max(s.size(-d) for s in group)
                    padded = []
                    for s in group:
                        pad = [0] * (pd * 2)
                        for d in range(1, pd + 1):
                            pad[2 * d - 1] = max_shape[d - 1] - s.size(-d)
                        padded.append(F.pad(s.data, pad, value=s.padding_value))
                    stacked.append(default_collate(padded))
                else:
                    stacked.append(default_collate([s.data for s in group]))
            return DataContainer(stacked, True, batch[0].padding_value)
        for i in range(0, len(batch), samples_per_gpu):
            stacked.append([s.data for s in batch[i:i + samples_per_gpu]])
        return DataContainer(stacked, batch[0].stack, batch[0].padding_value)
    if isinstance(batch[0], Sequence) and not isinstance(batch[0], (str, bytes)):
        return [collate(samples, samples_per_gpu) for samples in zip(*batch)]
    if isinstance(batch[0], Mapping):
        return {key: collate([d[key] for d in batch], samples_per_gpu) for key in batch[0]}
    return default_collate(batch)


MODULE_WRAPPERS = Registry("module wrapper")


def is_module_wrapper(module):
    return isinstance(module, tuple(MODULE_WRAPPERS.module_dict.values()))


@MODULE_WRAPPERS.register_module()
class MMDataParallel(nn.Module):
    """Single-process, single-device wrapper with mmcv's calling convention: `.module`, `device_ids`, keyword arguments
    scattered to device_ids[0] (DataContainers unwrapped) before the wrapped module is called (tools/test.py:105)."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__()
        self.module = module
        if device_ids is None:
            device_ids = [0] if torch.cuda.is_available() else []
        self.device_ids = list(device_ids)
        self.dim = dim
        if len(self.device_ids) > 1:
            raise NotImplementedError("one process drives one GPU (tools/test.py passes device_ids=[0])")
        if self.device_ids:
            self.module.to(torch.device("cuda", self.device_ids[0]))

    def scatter(self, inputs, kwargs, device_ids):
        return scatter_gather.scatter_kwargs(inputs, kwargs, device_ids, dim=self.dim)

    def forward(self, *inputs, **kwargs):
        inputs, kwargs = self.scatter(inputs, kwargs, self.device_ids if self.device_ids else [-1])
        return self.module(*inputs[0], **kwargs[0])

    def train_step(self, *inputs, **kwargs):
        inputs, kwargs = self.scatter(inputs, kwargs, self.device_ids if self.device_ids else [-1])
        return self.module.train_step(*inputs[0], **kwargs[0])

    def val_step(self, *inputs, **kwargs):
        inputs, kwargs = self.scatter(inputs, kwargs, self.device_ids if self.device_ids else [-1])
        return self.module.val_step(*inputs[0], **kwargs[0])


@MODULE_WRAPPERS.register_module()
class MMDistributedDataParallel(MMDataParallel):
    """One process per GPU (tools/test.py:108-111 with --launcher pytorch).  Inference needs no gradient
    synchronisation, so this is the single-device wrapper bound to the process's current device; the distributed
    sampler shards the dataset and mogen/apis/test.py collects the results."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0, broadcast_buffers=True,
                 find_unused_parameters=False, **kwargs):
        if device_ids is None and torch.cuda.is_available():
            device_ids = [torch.cuda.current_device()]
        super().__init__(module, device_ids=device_ids, output_device=output_device, dim=dim)
