"""Result hand-off: the oracle's restatement of scipy's Gaussian filter pinned against scipy
itself and against the arithmetic of tools/visualize.py:219-249 (CPU), the CUDA kernels against the oracle (GPU, bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import handoff_oracle as H


def _visualize_py_reference(pred, mean, std):
    """tools/visualize.py:219-249 with the library calls the tool itself makes (numpy + scipy.ndimage.gaussian_filter)."""
    from scipy.ndimage import gaussian_filter
    x = pred * std + mean
    T = x.shape[0]
    rec_pose = np.zeros((T, 165))
    rec_pose[:, :3 + 63] = x[:, :3 + 63]
    rec_pose[:, 66:66 + 3] = x[:, 66 + 90:66 + 93]
    rec_pose[:, 66 + 9:66 + 90 + 9] = x[:, 66:66 + 90]
    rec_trans = x[:, 309:309 + 3]
    rec_exp = x[:, 209:209 + 100]

    def filt(m, sigma):
        for i in range(m.shape[1]):
            m[:, i] = gaussian_filter(m[:, i], sigma=sigma, mode="nearest")
        return m
    return dict(poses=filt(rec_pose, 3.5), trans=filt(rec_trans, 3.0), expressions=filt(rec_exp, 2.0))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("T", [7, 120, 196])
def test_oracle_handoff_is_scipy_bit_for_bit(T, dtype):
    rng = np.random.default_rng(T)
    pred = rng.standard_normal((T, 322)).astype(np.float32)
    mean = rng.standard_normal(322).astype(dtype)
    std = (0.5 + rng.random(322)).astype(dtype)
    want = _visualize_py_reference(pred.copy(), mean, std)
    got = H.smplx_handoff(pred.copy(), mean, std)
    for k in want:
        assert got[k].dtype == want[k].dtype == (np.float64 if (k == "poses" or dtype == np.float64) else np.float32)
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    w, r = H.gaussian_weights(3.5)
    assert r == 14 and abs(w.sum() - 1.0) < 1e-15


def test_oracle_align_faces():
    rng = np.random.default_rng(1)
    pred, motion = rng.standard_normal((9, 322)).astype(np.float32), rng.standard_normal((9, 322)).astype(np.float32)
    out = H.align_faces(pred, motion)
    keep = np.r_[0:156, 309:312]
    np.testing.assert_array_equal(out[:, keep], pred[:, keep])
    np.testing.assert_array_equal(out[:, 156:309], motion[:, 156:309])
    np.testing.assert_array_equal(out[:, 312:], motion[:, 312:])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gpu_handoff_kernels_match_oracle_bit_for_bit(dtype):
    from motioncraft_b200 import handoff
    rng = np.random.default_rng(5)
    B, T = 3, 196
    lengths = [196, 64, 9]
    pred = rng.standard_normal((B, T, 322)).astype(np.float32)
    mean = rng.standard_normal(322).astype(dtype)
    std = (0.5 + rng.random(322)).astype(dtype)
    out = handoff.smplx_handoff(torch.from_numpy(pred).cuda(), mean, std, lengths=lengths)
    for b, n in enumerate(lengths):
        want = H.smplx_handoff(pred[b, :n].copy(), mean, std)
        for k in want:
            got = out[k][b].cpu().numpy()
            assert got.dtype == want[k].dtype
            np.testing.assert_array_equal(got[:n], want[k], err_msg=f"{k} sample {b}")
            assert not got[n:].any()
    # a long concatenated sequence as tools/visualize.py builds it for several prompts (one filter over the whole sequence)
    long_pred = rng.standard_normal((3000, 322)).astype(np.float32)
    got = handoff.smplx_handoff(torch.from_numpy(long_pred).cuda(), mean, std)
    want = H.smplx_handoff(long_pred.copy(), mean, std)
    for k in want:
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k], err_msg=k)
    motion = rng.standard_normal((B, T, 322)).astype(np.float32)
    p = torch.from_numpy(pred.copy()).cuda()
    handoff.align_faces_(p, torch.from_numpy(motion).cuda())
    np.testing.assert_array_equal(p.cpu().numpy().reshape(-1, 322), H.align_faces(pred.reshape(-1, 322), motion.reshape(-1, 322)))
