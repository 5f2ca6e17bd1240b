"""Condition pre-encoders of the ControlNet branch (mogen/models/transformers/controlnet_mcm.py:88-104,
mogen/models/utils/blocks.py:11-71).

The speech-to-gesture model feeds raw 16 kHz audio (B, samples, channels) through a strided residual Conv1d stack
(`WavEncoder`) before `control_cond_input`.  The reference re-runs it on every denoise step although it depends only on
the condition (`forward_c`, controlnet_mcm.py:155-166); here it is evaluated ONCE per sampling run, outside the per-step
path, with cuDNN convolutions through torch (step-invariant, library convolution).
Parameter names and shapes are the reference's, so `condition_pre_encoder.pre_encoder.feat_extractor.{i}.*` checkpoint
entries load unchanged.  Inference only: BatchNorm uses its running statistics.
"""
import torch
import torch.nn.functional as F
from torch import nn

from ._lib import McmError

_KERNEL = 15
# (out channels as a fraction of out_dim, stride, padding of the strided convolutions, has projection shortcut)
_STAGES = ((4, 5, 1600, True), (4, 6, 0, True), (4, 1, 7, False), (2, 6, 0, True), (2, 1, 7, False), (1, 3, 0, True))


class BasicBlock(nn.Module):
    """Conv1d(k=15, stride s) - BN - LeakyReLU - Conv1d(k=15, 'same') - BN, plus an identity or strided-conv shortcut,
    then LeakyReLU (blocks.py:11-51)."""

    def __init__(self, inplanes, planes, stride, first_padding, project):
        super().__init__()
        self.stride, self.first_padding = stride, first_padding
        self.conv1 = nn.Conv1d(inplanes, planes, _KERNEL, stride=stride, padding=first_padding)
        self.bn1 = nn.BatchNorm1d(planes)
        self.conv2 = nn.Conv1d(planes, planes, _KERNEL, padding=_KERNEL // 2)
        self.bn2 = nn.BatchNorm1d(planes)
        self.downsample = nn.Sequential(nn.Conv1d(inplanes, planes, _KERNEL, stride=stride, padding=first_padding),
                                        nn.BatchNorm1d(planes)) if project else None

    @staticmethod
    def _bn(x, bn):
        return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)

    def forward(self, x):
        y = F.leaky_relu(self._bn(self.conv1(x), self.bn1), 0.01)
        y = self._bn(self.conv2(y), self.bn2)
        skip = x if self.downsample is None else self._bn(self.downsample[0](x), self.downsample[1])
        return F.leaky_relu(y + skip, 0.01)


class WavEncoder(nn.Module):
    """(B, samples[, channels]) -> (B, frames, out_dim); 159 900 samples -> 297 frames (blocks.py:53-71)."""

    def __init__(self, out_dim, audio_in=1):
        super().__init__()
        self.out_dim = out_dim
        blocks, cin = [], audio_in
        for frac, stride, pad, project in _STAGES:
            blocks.append(BasicBlock(cin, out_dim // frac, stride, pad, project))
            cin = out_dim // frac
        self.feat_extractor = nn.Sequential(*blocks)

    def forward(self, wav):
        wav = wav.unsqueeze(1) if wav.dim() == 2 else wav.transpose(1, 2)
        return self.feat_extractor(wav).transpose(1, 2)


class ConditionEncoder(nn.Module):
    """controlnet_mcm.py:88-104: only the BEAT2 raw-waveform encoder exists in the reference."""

    def __init__(self, condition_encode_cfg):
        super().__init__()
        get = (lambda k: condition_encode_cfg[k]) if hasattr(condition_encode_cfg, "__getitem__") else \
              (lambda k: getattr(condition_encode_cfg, k))
        if get("dataset_name") != "beats2" or get("condition_pre_encode_type") != "wav":
            raise McmError("the reference implements condition pre-encoding for dataset_name='beats2', type 'wav' only")
        self.pre_encoder = WavEncoder(out_dim=get("condition_latent_dim"), audio_in=get("control_cond_feats"))
        self.raw_feats = get("control_cond_feats")

    @torch.no_grad()
    def forward(self, condition):
        return self.pre_encoder(condition)
