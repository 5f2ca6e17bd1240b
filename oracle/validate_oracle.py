"""TEST INFRASTRUCTURE ONLY -- check oracle/mcm_oracle.py against the UNMODIFIED reference
(imported from /root/reference under oracle/ref_shim.py).  Build-container only.

usage: python oracle/validate_oracle.py [--T 60] [--B 2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mcm_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from motioncraft_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=60)
    ap.add_argument("--B", type=int, default=2)
    args = ap.parse_args()
    torch.manual_seed(0)
    T, B = args.T, args.B
    ref = ref_shim.build_reference_mcm(T=T)
    sd = synth.synth_state_dict({k: v.shape for k, v in ref.state_dict().items()})
    ref.load_state_dict(sd)
    x = synth.synth_tensor("x_T", (B, T, 322), synth.SEED_XT)
    xf_out = synth.synth_tensor("xf_out", (B, 77, 256), synth.SEED_XF_OUT)
    xf_proj = synth.synth_tensor("xf_proj", (B, 2048), synth.SEED_XF_PROJ)
    t = torch.full((B,), 999, dtype=torch.long)
    with torch.no_grad():
        want = ref(x, t, motion_mask=torch.ones(B, T), motion_length=torch.full((B,), T),
                   xf_proj=xf_proj, xf_out=xf_out)
        got = O.mcm_forward(sd, x, t, xf_proj, xf_out)
    print("forward max|diff| =", (want - got).abs().max().item(), " |eps|max =", want.abs().max().item(),
          "rms =", want.pow(2).mean().sqrt().item())
    assert torch.equal(want, got), "oracle forward is not bit-identical to the reference"

    # DDIM-50
    diff = ref_shim.build_reference_diffusion("15,15,8,6,6")
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    assert tmap == diff.timestep_map
    kw = dict(motion_mask=torch.ones(B, T), motion_length=torch.full((B,), T), xf_proj=xf_proj,
              xf_out=xf_out, y={})
    with torch.no_grad():
        want = diff.ddim_sample_loop(ref, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                     model_kwargs=kw, eta=0)
        got = O.ddim_sample_loop(lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out), x.clone(),
                                 tables, tmap)
    print("ddim50 max|diff| =", (want - got).abs().max().item(), "|x0|max =", want.abs().max().item())
    assert torch.equal(want, got), "oracle DDIM loop is not bit-identical to the reference"

    # RePaint / outpainting (harmonising loop + plain loop), scripted randn_like
    import copy
    from oracle.make_golden import scripted_randn_like
    from mogen.models.utils.scheduler import get_schedule_jump_cjm_ddim
    for args in [(25, 1, 1), (50, 3, 5), (50, 1, 1), (50, 2, 3), (100, 4, 2)]:
        assert get_schedule_jump_cjm_ddim(*args) == O.schedule_jump_cjm_ddim(*args), args
    L = 10
    gt = torch.zeros(T, 322)
    mask = torch.zeros(T, 322, dtype=torch.bool)
    gt[:L] = synth.synth_tensor("gt", (T, 322), synth.SEED_REPAINT_GT)[:L]
    mask[:L] = True
    for mode in ("harmonize", "plain"):
        d = ref_shim.build_reference_diffusion("15,15,8,6,6")
        d.opt = copy.copy(d.opt)
        d.opt.overlap_len, d.opt.no_repaint = L, (mode == "plain")
        times = None if mode == "plain" else O.schedule_jump_cjm_ddim(50, d.opt.jump_length, d.opt.jump_n_sample)
        n_den = 50 if times is None else sum(1 for a_, b_ in zip(times[:-1], times[1:]) if b_ < a_)
        n_draw = 2 * n_den + (0 if times is None else len(times) - 1 - n_den)
        noise = synth.synth_tensor("repaint_noise", (n_draw, B, T, 322), synth.SEED_REPAINT_NOISE)
        with torch.no_grad(), scripted_randn_like([noise[i] for i in range(n_draw)]):
            want = d.ddim_sample_loop(ref, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                      model_kwargs=dict(kw, y={"gt": gt.clone(), "outpainting_mask": mask}), eta=0)
        with torch.no_grad():
            got = O.ddim_repaint_loop(lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out), x.clone(), tables, tmap,
                                      tables["betas"], gt, mask, [noise[i] for i in range(n_draw)], times=times, overlap_len=L)
        print(f"repaint {mode}: max|diff| =", (want - got).abs().max().item())
        assert torch.equal(want, got), f"oracle RePaint loop ({mode}) is not bit-identical to the reference"
    print("OK")


if __name__ == "__main__":
    main()
