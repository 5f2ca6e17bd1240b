"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the result hand-off of the reference:

  * tools/visualize.py:219-263 (motionx branch): de-normalise `pred * std + mean`, repack the 322-dim vector into
    SMPL-X poses (165) / expressions (100) / translation (3), Gaussian temporal filter per column
    (`scipy.ndimage.gaussian_filter(col, sigma, mode="nearest")`, sigma 3.5 / 2.0 / 3.0);
  * mogen/datasets/base_dataset.py:121-125: the evaluator overwrites the face / shape columns of the prediction with the
    ground truth.

The filter is a third-party dependency of the reference (scipy; `requirements.txt` leaves it unpinned, scipy 1.18.1 is
installed here): `gaussian_filter1d` below restates its published algorithm -- weights of `_gaussian_kernel1d`
(scipy/ndimage/_filters.py) and the symmetric branch of `NI_Correlate1D` (scipy/ndimage/src/ni_filters.c: centre tap first,
then `tmp += (x[l + j] + x[l - j]) * w[j]` from the farthest tap inwards, edge samples replicated) -- and is pinned bit for
bit against scipy itself in tests/test_handoff.py.  Only tests/ may import this file.
"""
import numpy as np

POSE_SIGMA, TRANS_SIGMA, EXPR_SIGMA = 3.5, 3.0, 2.0          # tools/visualize.py:247-249


def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(truncate * sigma + 0.5)."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum(), lw


def gaussian_filter1d(col, sigma):
    """gaussian_filter(col, sigma, mode='nearest') of a 1-D array, in scipy's own accumulation order: the line is converted
    to double, accumulated in double, and the result cast to the array's dtype (float32 columns stay float32)."""
    dtype = col.dtype
    w, lw = gaussian_weights(sigma)
    n = col.shape[0]
    ext = np.concatenate([np.full(lw, col[0]), col, np.full(lw, col[-1])]).astype(np.float64)
    out = np.empty(n, dtype=np.float64)
    c = lw                                        # centre tap of the (symmetric) kernel
    for l in range(n):
        p = l + lw
        tmp = ext[p] * w[c]
        for j in range(-lw, 0):
            tmp += (ext[p + j] + ext[p - j]) * w[c + j]
        out[l] = tmp
    return out.astype(dtype)


def smplx_handoff(pred, mean, std):
    """pred (T, 322) float32 -> dict(poses (T, 165) float64, expressions (T, 100), trans (T, 3)): the latter two are slices
    of the de-normalised array and keep ITS dtype (float32 when mean / std are float32).   visualize.py:219-249"""
    x = pred * std + mean                         # numpy promotion rules of the reference apply (float32 or float64)
    T = x.shape[0]
    pose = np.zeros((T, 165))
    pose[:, :3 + 63] = x[:, :3 + 63]
    pose[:, 66:66 + 3] = x[:, 66 + 90:66 + 93]
    pose[:, 66 + 9:66 + 90 + 9] = x[:, 66:66 + 90]
    trans = x[:, 309:309 + 3].copy()
    expr = x[:, 209:209 + 100].copy()
    for arr, sigma in ((pose, POSE_SIGMA), (trans, TRANS_SIGMA), (expr, EXPR_SIGMA)):
        for i in range(arr.shape[1]):
            arr[:, i] = gaussian_filter1d(arr[:, i], sigma)
    return dict(poses=pose, expressions=expr, trans=trans)


def align_faces(pred, motion):
    """base_dataset.py:121-125 (in place on a copy)."""
    out = pred.copy()
    out[:, 156:309] = motion[:, 156:309]
    out[:, 312:] = motion[:, 312:]
    return out
