"""Developer probe for compute-sanitizer: one decoder layer at the shapes that select every kernel variant
(fused / kernel-per-op, T = 196 / 300 / 1024, control branch packing), small batches.

    compute-sanitizer --tool memcheck python tools/memcheck_probe.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import modules, synth
from motioncraft_b200.engine import DenoiserEngine

for T, B, fused in ((196, 12, 1), (196, 2, 0), (300, 2, 0), (1024, 2, 0), (1024, 3, 1)):
    sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1)).items() if ".ffn_channel." not in k}
    eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=1)
    eng.set_option("graph", 0)
    if not fused:
        eng.set_option("fused", 0); eng.set_option("fused_sa", 0)
    else:
        eng.set_option("fused_min_rows", 0)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, 322, generator=g).cuda()
    eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
    out = eng.denoise(x, 500)
    torch.cuda.synchronize()
    print(T, B, fused, bool(torch.isfinite(out).all()), flush=True)
    eng.close()

# Path-B pieces (SFFN, STMA after its MoE layers)
from motioncraft_b200 import pathb
sf = pathb.SFFN(latent_dim=64, ffn_dim=128, dropout=0.0, time_embed_dim=256, num_heads=12).cuda()
print("sffn", bool(torch.isfinite(sf(torch.randn(2, 9, 768).cuda(), torch.randn(2, 256).cuda())).all()), flush=True)
for L, dyn, Ht in ((64, True, 1), (32, False, 12)):
    tm = pathb.STMATail(latent_dim=L, num_heads=12, num_text_heads=Ht, time_embed_dim=256, dynamic_body=dyn).cuda()
    o = tm(torch.randn(3, 20, 12 * L).cuda(), torch.randn(3, 20, 12, 4 * L).cuda(), torch.randn(3, 7, Ht, 2 * L).cuda(),
           torch.randn(3, 256).cuda(), torch.ones(3, 20, 1).cuda(), torch.tensor([1, 0, 11]).view(3, 1, 1).cuda())
    print("stma", L, dyn, bool(torch.isfinite(o).all()), flush=True)
print("done")
