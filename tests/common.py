"""Shared helpers for the parity tests: seeded synthetic weights / inputs and the oracle runs."""
import functools

import numpy as np
import torch

from motioncraft_b200 import modules, synth
from oracle import mcm_oracle as O


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def max_rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@functools.lru_cache(maxsize=4)
def base_state(T, num_layers=8):
    sd = synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=num_layers))
    return sd


def hot(sd):
    return {k: v for k, v in sd.items() if ".ffn_channel." not in k}


def inputs(B, T, n_tokens=77):
    x = synth.synth_tensor("x_T", (B, T, 322), synth.SEED_XT)
    xf_out = synth.synth_tensor("xf_out", (B, n_tokens, 256), synth.SEED_XF_OUT)
    xf_proj = synth.synth_tensor("xf_proj", (B, 2048), synth.SEED_XF_PROJ)
    return x, xf_out, xf_proj


def oracle_forward(sd, x, t, xf_proj, xf_out, dtype=torch.float32, collect=None):
    sd_ = {k: v.to(dtype) for k, v in sd.items()}
    tt = torch.full((x.shape[0],), t, dtype=torch.long) if isinstance(t, int) else t
    with torch.no_grad():
        return O.mcm_forward(sd_, x.to(dtype), tt, xf_proj.to(dtype), xf_out.to(dtype), collect=collect)


def oracle_ddim(sd, x, xf_proj, xf_out, respace="15,15,8,6,6", dtype=torch.float32):
    sd_ = {k: v.to(dtype) for k, v in sd.items()}
    tables, tmap = O.spaced_tables(1000, respace)
    with torch.no_grad():
        return O.ddim_sample_loop(lambda xx, tt: O.mcm_forward(sd_, xx, tt, xf_proj.to(dtype), xf_out.to(dtype)),
                                  x.to(dtype), tables, tmap)


def oracle_ddpm(sd, x, xf_proj, xf_out, step_noise, respace="10", dtype=torch.float32):
    sd_ = {k: v.to(dtype) for k, v in sd.items()}
    tables, tmap = O.spaced_tables(1000, respace)
    with torch.no_grad():
        return O.p_sample_loop(lambda xx, tt: O.mcm_forward(sd_, xx, tt, xf_proj.to(dtype), xf_out.to(dtype)),
                               x.to(dtype), tables, tmap, step_noise.to(dtype))


ctrl_shapes = modules.ctrl_state_shapes
engine_state_from_ctrl = modules.engine_state_from_ctrl
