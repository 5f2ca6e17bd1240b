"""Developer probe (GPU box): stage-by-stage check of the fused cross-attention + FFN kernel.

For each truncation point k ("fused_stop") the kernel dumps its shared-memory operand tile; this script compares
the dump (and h / the hidden scratch where they are the phase's product) with a float64 torch evaluation of the
reference formulas, then compares the complete fused block and a full 50-step run with the unfused path.
usage: python tools/fused_check.py [T] [B]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import modules, synth  # noqa: E402
from motioncraft_b200.engine import DenoiserEngine, SamplerTables  # noqa: E402
from oracle import mcm_oracle as O  # noqa: E402
import torch.nn.functional as Fn  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 196
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.manual_seed(0)
    sd = synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1))
    sd = {k: v for k, v in sd.items() if ".ffn_channel." not in k}
    g = torch.Generator().manual_seed(5)
    h0 = torch.randn(B, T, 512, generator=g)
    emb = torch.randn(B, 2048, generator=g)
    xf_out = torch.randn(B, 77, 256, generator=g)
    xf_proj = torch.randn(B, 2048, generator=g)
    eng = DenoiserEngine(sd, seq_len=T, num_layers=1, max_batch=B)
    eng.prepare_conditions(xf_out, xf_proj)
    eng.set_option("fused_min_rows", 0)
    eng.set_option("dual", 0)
    eng.set_option("graph", 0)

    from tests import common as C
    st = C.block_stages(sd, h0, emb, xf_out)
    want = {k: st[k] for k in (1, 2, 3, 4, 6)}
    h_ca, hid, h_out = st[4], st["hid"], st["out"]

    eng.set_option("fused", 0)
    ref_block = eng.block_forward(0, 0, h0, emb)
    print(f"unfused block vs fp64: {rel(ref_block, h_out):.3e}", flush=True)
    eng.set_option("fused", 1)
    n = B * T * 512
    for k in (1, 2, 3, 4, 5, 6):
        eng.set_option("fused_stop", k)
        hk = eng.block_forward(0, 0, h0, emb)
        torch.cuda.synchronize()
        dump = eng.debug_copy(0, n).view(B, T, 512).float()
        msg = f"stop={k}:"
        if k in want:
            msg += f" operand tile rel = {rel(dump, want[k]):.3e}"
        if k >= 4:
            msg += f"  h rel = {rel(hk, h_ca):.3e}"
        if k == 5:
            hs = eng.debug_copy(1, B * T * 1024)
            rows = B * T
            hs = hs.view(-1, 1024)[:rows].view(B, T, 1024).float()
            msg += f"  hidden rel = {rel(hs, hid):.3e} (valid when tiles <= pairs)"
        print(msg, flush=True)
    eng.set_option("fused_stop", 0)
    out = eng.block_forward(0, 0, h0, emb)
    print(f"fused block vs fp64: {rel(out, h_out):.3e}   vs unfused: {rel(out, ref_block):.3e}", flush=True)
    for lvl in (0, 1, 2, 3):
        eng.set_option("fused_sa", lvl)
        o = eng.block_forward(0, 0, h0, emb)
        print(f"fused CA/FFN, channel-attention fusion level {lvl} vs fp64: {rel(o, h_out):.3e}", flush=True)
    eng.set_option("fused_sa", 1)
    out = eng.block_forward(0, 0, h0, emb)
    out2 = eng.block_forward(0, 0, h0, emb)
    print("fused deterministic:", torch.equal(out, out2), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
