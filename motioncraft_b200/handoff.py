"""Result hand-off on the device -- Python side of `mcm_handoff_*` (include/mcm_b200.h).

What the reference does on the host after sampling, before files / evaluators see the motion:
  * `smplx_handoff`  tools/visualize.py:219-263 (motionx): de-normalise, 322 -> SMPL-X poses / expressions / trans, Gaussian
    temporal filter per column (scipy.ndimage.gaussian_filter, mode="nearest");
  * `align_faces_`   mogen/datasets/base_dataset.py:121-125: face / shape columns of the prediction := ground truth.
Both run as CUDA kernels on the sampler's output, so the only device->host traffic left is the final result.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import McmError

POSE_SIGMA, TRANS_SIGMA, EXPR_SIGMA = 3.5, 3.0, 2.0      # tools/visualize.py:247-249


def gaussian_taps(sigma, truncate=4.0):
    """Normalised taps of scipy's `_gaussian_kernel1d(sigma, 0, radius)`, radius = int(truncate * sigma + 0.5) -- the same
    numpy expressions, so the float64 values are the ones scipy uses."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum(), lw


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def smplx_handoff(pred, mean=None, std=None, lengths=None):
    """pred (B, T, 322) or (T, 322) fp32 CUDA tensor -> dict(poses (.., 165), expressions (.., 100), trans (.., 3)),
    float64 CUDA tensors, bit-identical to tools/visualize.py:219-249 run on the host.  mean / std: numpy arrays as the
    tool loads them (float32 or float64; default zeros / ones as visualize.py:190-193); lengths: valid frames per sample."""
    if pred.device.type != "cuda":
        raise McmError("motioncraft_b200 hand-off kernels run on the CUDA device that holds the sampler output")
    squeeze = pred.dim() == 2
    x = pred.detach().to(torch.float32).contiguous()
    if squeeze:
        x = x.unsqueeze(0)
    B, T, Fd = x.shape
    if Fd != 322:
        raise McmError("the SMPL-X repack is defined for the 322-dim motionx vector (tools/visualize.py:241-246)")
    mean = np.zeros(322) if mean is None else np.asarray(mean)
    std = np.ones(322) if std is None else np.asarray(std)
    f32 = int(mean.dtype == np.float32 and std.dtype == np.float32)      # numpy: float32 * float32 + float32 stays float32
    dev = x.device
    mean_d = torch.from_numpy(mean.astype(np.float64)).to(dev)
    std_d = torch.from_numpy(std.astype(np.float64)).to(dev)
    taps = [gaussian_taps(s) for s in (POSE_SIGMA, EXPR_SIGMA, TRANS_SIGMA)]
    w = [torch.from_numpy(np.ascontiguousarray(t[0])).to(dev) for t in taps]
    len_d = None
    if lengths is not None:
        len_d = torch.as_tensor(lengths).to(device=dev, dtype=torch.int32).contiguous()
    pose = torch.empty(B, T, 165, device=dev, dtype=torch.float64)
    expr = torch.empty(B, T, 100, device=dev, dtype=torch.float64)
    trans = torch.empty(B, T, 3, device=dev, dtype=torch.float64)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mcm_handoff_smplx(_ptr(x), B, T, _ptr(len_d), _ptr(mean_d), _ptr(std_d), f32, _ptr(w[0]), taps[0][1],
                                         _ptr(w[1]), taps[1][1], _ptr(w[2]), taps[2][1], _ptr(pose), _ptr(expr), _ptr(trans),
                                         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.current_stream(dev).synchronize()      # the staged parameter tensors go out of scope
    if f32:            # the reference's expressions / trans are float32 slices then (values already rounded by the kernel)
        expr, trans = expr.float(), trans.float()
    out = dict(poses=pose, expressions=expr, trans=trans)
    return {k: v[0] for k, v in out.items()} if squeeze else out


class Denormaliser:
    """`pred * std + mean` on the device with numpy's promotion rules (float32 when both arrays are float32, else float64);
    the statistics are uploaded once."""

    def __init__(self, mean, std, device):
        mean, std = np.asarray(mean), np.asarray(std)
        self.f32 = int(mean.dtype == np.float32 and std.dtype == np.float32)
        self.device = torch.device(device)
        self.mean = torch.from_numpy(mean.astype(np.float64)).to(self.device)
        self.std = torch.from_numpy(std.astype(np.float64)).to(self.device)

    def __call__(self, pred, want64=True, want32=False):
        """pred (..., F) fp32 CUDA tensor -> (float64 result or None, its float32 rounding or None)."""
        x = pred.detach().to(torch.float32).contiguous()
        Fd = x.shape[-1]
        o64 = torch.empty(x.shape, device=x.device, dtype=torch.float64) if want64 else None
        o32 = torch.empty(x.shape, device=x.device, dtype=torch.float32) if want32 else None
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.mcm_handoff_denorm(_ptr(x), _ptr(self.mean), _ptr(self.std), x.numel() // Fd, Fd, self.f32, _ptr(o64),
                                              _ptr(o32), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return o64, o32


def align_faces_(pred, motion):
    """In place on the device: pred[..., 156:309] = motion[..., 156:309]; pred[..., 312:] = motion[..., 312:]
    (base_dataset.py:121-125).  pred, motion: (..., 322) fp32 CUDA tensors of the same shape; returns pred."""
    if pred.device.type != "cuda" or motion.device != pred.device:
        raise McmError("align_faces_ works on CUDA tensors on one device")
    if pred.shape != motion.shape or pred.shape[-1] != 322 or pred.dtype != torch.float32 or not pred.is_contiguous():
        raise McmError("align_faces_ needs contiguous fp32 (..., 322) tensors of equal shape")
    m = motion.detach().to(torch.float32).contiguous()
    lib = _lib.load()
    with torch.cuda.device(pred.device):
        _lib.check(lib.mcm_handoff_align_faces(_ptr(pred), _ptr(m), pred.numel() // 322, 322,
                                               ctypes.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)))
    return pred
