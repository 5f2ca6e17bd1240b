"""Developer probe: per-GEMM-group device time of one denoise step via the library's event timing."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from motioncraft_b200 import _lib, modules, synth
from motioncraft_b200.engine import DenoiserEngine

name = sys.argv[1] if len(sys.argv) > 1 else "t2m"
wl = dict(WORKLOADS[name])
if len(sys.argv) > 2:
    wl["B"] = int(sys.argv[2])
B, T = wl["B"], wl["T"]
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
eng = DenoiserEngine(sd, seq_len=T, max_batch=B)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, 322, generator=g).cuda()
eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
for _ in range(3):
    eng.denoise(x, 500)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.denoise(x, 500)
e1.record()
torch.cuda.synchronize()
print(f"MCM_DEBUG_EPI={os.environ.get('MCM_DEBUG_EPI','0')}: denoise step {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
_lib.timing_enable(True)
eng.denoise(x, 500)
tm = _lib.timing_collect()
_lib.timing_enable(False)
print("   gemm ms", round(tm['gemm']['ms'], 3), "row ms", round(tm['row']['ms'], 3), "fused ms", round(tm['fused']['ms'], 3),
      "fused TFLOP/s", round(tm['fused']['flops'] / max(tm['fused']['ms'], 1e-9) / 1e9, 1))
