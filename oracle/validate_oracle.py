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
    print("OK")


if __name__ == "__main__":
    main()
