"""Developer probe: wall-clock latency of one 50-step DDIM run at small batch (launch-bound regime)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import modules, synth
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from motioncraft_b200.diffusion import build_diffusion
T = 196
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
d = build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small", respace="15,15,8,6,6"))
st = SamplerTables(d._tables(), d.timestep_map, "ddim", 0.0)
for B in (1, 8, 32):
    eng = DenoiserEngine(sd, seq_len=T, max_batch=B)
    for kv in sys.argv[1:]:                       # e.g. fused_sa_min_rows=0
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    x = torch.randn(B, T, 322).cuda()
    eng.prepare_conditions(torch.randn(B, 77, 256).cuda(), torch.randn(B, 2048).cuda())
    for _ in range(3):
        eng.sample(st, x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        eng.sample(st, x)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"B={B:3d} graph={os.environ.get('MCM_GRAPH', '1')}: {ms:8.2f} ms per 50-step run  ({B * T / ms * 1e3:9.0f} frames/s)", flush=True)
    eng.close()
