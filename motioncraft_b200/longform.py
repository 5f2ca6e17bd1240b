"""Long-form generation: the sliding-window drivers of tools/m2d_test.py:145-222 and tools/s2g_test.py:144-241 as ONE
device-side pipeline.

The reference generates a long sequence window by window on the HOST: every window is a complete sampling run
(`model(**input)` -> `.cpu().numpy()` -> `pred * std + mean` -> `torch.tensor(...)` back to the GPU), window i + 1 pins its
first `overlap_len` frames to the tail of window i through y['gt'] / y['outpainting_mask'] (RePaint), and songs / speeches
are processed one after the other.  Here

  * the windows of MANY sequences form the batch (they are independent; only window i -> i + 1 of one sequence is a
    dependency), so the denoiser runs at B = number of sequences instead of B = 1;
  * the hand-over stays on the device: `mcm_sample_repaint` writes x_0, `mcm_handoff_denorm` de-normalises it, the tail
    becomes the next window's y['gt'], the kept part is copied into the result -- all enqueued on one stream;
  * nothing synchronises with the host between windows: the whole pipeline is enqueued, the caller synchronises once.

Semantics kept from the tools (including their quirk that the pinned frames are the DE-NORMALISED tail of the previous
window, :189 / :201 in m2d_test.py, :205 / :218 in s2g_test.py):
    round_l = motion_length - pre_frames;  window i covers frames [i round_l, i round_l + motion_length)
    window 0: mask empty (or the first overlap_len frames of `first_gt` when fix_very_first) -> plain / harmonising loop
    window i > 0: mask[:overlap_len] = True, gt[:overlap_len] = outputs_{i-1}[-overlap_len:]
    result (repaint): concat(window_i[:round_l] for i < last, window_last)
"""
import torch

from ._lib import McmError
from .engine import SamplerTables
from .handoff import Denormaliser
from .scheduler import count_draws, get_schedule_jump_cjm_ddim


def window_count(total_len, motion_length, pre_frames):
    """roundt / round_l of tools/m2d_test.py:143-145."""
    round_l = motion_length - pre_frames
    return (total_len - pre_frames) // round_l, round_l


def sample_windows(model, diffusion, window_kwargs, *, motion_length, pre_frames, mean, std, opt=None, first_gt=None,
                   noise=None, repaint_noise=None, input_feats=322):
    """Generate `len(window_kwargs)` overlapping windows for a batch of sequences.

    model          MCMTransformer / ControlT2MHalf_MCM (eval, on the CUDA device)
    diffusion      SpacedDiffusion built with the tools' `opt` namespace (overlap_len, addBlend, no_repaint, jump_*, ...)
    window_kwargs  per window: the model kwargs of that window for ALL sequences (xf_proj (B, E), xf_out (B, N, L) or text /
                   clip_feat, c (B, len, feats)) -- tools/m2d_test.py:166-175
    first_gt       (B, overlap_len, F) ground truth of the very first frames (--fix_very_first) or None
    noise          optional list of x_T per window (B, motion_length, F); repaint_noise: optional list of draw tensors
    returns        (B, (W - 1) * round_l + motion_length, F) de-normalised motion, float64 (float32 when mean / std are
                   float32, as numpy would give), on the device.  Nothing in here waits for the GPU.
    """
    opt = opt if opt is not None else diffusion.opt
    if opt is None:
        raise McmError("long-form sampling needs the tools' `opt` namespace (overlap_len, addBlend, no_repaint, ...)")
    if getattr(opt, "same_overlap_noisy", False):
        raise McmError("opt.same_overlap_noisy is dead code in the reference (gaussian_diffusion.py:881 never creates "
                       "saved_noisy_tail); not implemented")
    overlap = int(getattr(opt, "overlap_len", 0))
    W = len(window_kwargs)
    if W < 1 or not (0 <= overlap <= pre_frames < motion_length):
        raise McmError("need at least one window and 0 <= overlap_len <= pre_frames < motion_length")
    round_l = motion_length - pre_frames
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only")
    denorm = Denormaliser(mean, std, dev)
    times = None
    if not getattr(opt, "no_repaint", False):
        n = int(str(opt.timestep_respacing)[4:])
        times = (get_schedule_jump_cjm_ddim(n) if getattr(opt, "no_resample", False)
                 else get_schedule_jump_cjm_ddim(n, jump_length=opt.jump_length, jump_n_sample=opt.jump_n_sample))
    n_draws = count_draws(times, diffusion.num_timesteps)
    out_dtype = torch.float32 if denorm.f32 else torch.float64
    result, prev32 = None, None
    for i, kw in enumerate(window_kwargs):
        kw = dict(kw)
        B = (kw["xf_out"] if kw.get("xf_out") is not None else kw["c"]).shape[0]
        shape = (B, motion_length, input_feats)
        x_T = noise[i].to(dev) if noise is not None else torch.randn(*shape, device=dev)
        eng = model.bind_for_sampling(B, kw, dev)
        masked = overlap > 0 and (i > 0 or first_gt is not None)
        if masked:
            gt = torch.zeros(shape, device=dev)
            keep = torch.zeros(shape, device=dev, dtype=torch.bool)
            keep[:, :overlap] = True
            gt[:, :overlap] = prev32[:, -overlap:] if i > 0 else first_gt.to(dev, torch.float32)[:, :overlap]
            rn = repaint_noise[i].to(dev) if repaint_noise is not None else None
            seed = 0 if rn is not None else diffusion._draw_seed_async(i)
            tables = SamplerTables(diffusion._tables(), diffusion.timestep_map, "ddim", 0.0, seed=seed)
            x0 = eng.sample_repaint(tables, x_T, gt, keep, rn, times=times, betas=diffusion.betas, overlap_len=overlap,
                                    add_blend=bool(getattr(opt, "addBlend", True)))
        else:
            # no frame is pinned: `True in mask` is false and ddim_sample_loop takes the plain loop (gaussian_diffusion.py:962)
            tables = SamplerTables(diffusion._tables(), diffusion.timestep_map, "ddim", 0.0)
            x0 = eng.sample(tables, x_T)
        out64, prev32 = denorm(x0, want64=True, want32=True)       # `outputs` of the tools (:203-205), kept on the device
        if result is None:
            result = torch.empty(B, (W - 1) * round_l + motion_length, input_feats, device=dev, dtype=out_dtype)
        piece = out64 if i == W - 1 else out64[:, :round_l]
        result[:, i * round_l: i * round_l + piece.shape[1]] = piece.to(out_dtype)
    assert n_draws >= 0
    return result
