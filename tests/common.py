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


def block_stages(sd, h0, emb, xf_out, layer=0, dtype=torch.float64):
    """One DecoderLayer (mcm.py:25-41) in float64, returning the intermediate tensors the fused cross-attention + FFN
    kernel materialises, keyed by its debug truncation point (`fused_stop`):
      1: LN(h)   2: exp(q - max) per head (the kernel defers the softmax normalisation)   3: SiLU(AdaLN(LN(q ctx)))
      4: h after the cross-attention residual   'hid': GELU(linear1)   6: SiLU(AdaLN(LN(linear2)))   'out': layer output
    """
    import torch.nn.functional as Fn
    P = f"temporal_decoder_blocks.{layer}."
    d = {k: v.to(dtype) for k, v in sd.items() if k.startswith(P)}
    h0, emb, xf_out = h0.to(dtype), emb.to(dtype), xf_out.to(dtype)
    B, T, D = h0.shape
    with torch.no_grad():
        h_sa = O.efficient_self_attention(h0.transpose(1, 2), emb, d, P + "sa_block", 4).transpose(1, 2).contiguous()
        ca = P + "ca_block"
        ln = Fn.layer_norm(h_sa, (D,), d[ca + ".norm.weight"], d[ca + ".norm.bias"])
        q = Fn.linear(ln, d[ca + ".query.weight"], d[ca + ".query.bias"]).view(B, T, 4, D // 4)
        qe = torch.exp(q - q.max(dim=-1, keepdim=True).values).reshape(B, T, D)
        qs = torch.softmax(q, dim=-1)
        ctx = O.cross_attention_context(xf_out, d, ca, 4)
        y = torch.einsum("bnhd,bhdl->bnhl", qs, ctx).reshape(B, T, D)

        def styl(yy, pfx):
            eo = Fn.linear(Fn.silu(emb), d[pfx + ".emb_layers.1.weight"], d[pfx + ".emb_layers.1.bias"]).unsqueeze(1)
            sc, sh = eo.chunk(2, dim=2)
            return Fn.silu(Fn.layer_norm(yy, (D,), d[pfx + ".norm.weight"], d[pfx + ".norm.bias"]) * (1 + sc) + sh)

        a2 = styl(y, ca + ".proj_out")
        h_ca = h_sa + Fn.linear(a2, d[ca + ".proj_out.out_layers.2.weight"], d[ca + ".proj_out.out_layers.2.bias"])
        fn = P + "ffn_temporal"
        hid = Fn.gelu(Fn.linear(h_ca, d[fn + ".linear1.weight"], d[fn + ".linear1.bias"]))
        y2 = Fn.linear(hid, d[fn + ".linear2.weight"], d[fn + ".linear2.bias"])
        a5 = styl(y2, fn + ".proj_out")
        out = h_ca + Fn.linear(a5, d[fn + ".proj_out.out_layers.2.weight"], d[fn + ".proj_out.out_layers.2.bias"])
    return {1: ln, 2: qe, 3: a2, 4: h_ca, "hid": hid, 6: a5, "out": out}
