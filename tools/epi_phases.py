"""Developer probe: cycles the GEMM epilogue warps spend per phase, per GEMM of a decoder layer (MCM_DEBUG_EPI=3)."""
import ctypes, os, sys
os.environ["MCM_DEBUG_EPI"] = "3"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import _lib, modules, synth
from motioncraft_b200.engine import DenoiserEngine
B, T = 256, 196
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1)).items() if ".ffn_channel." not in k}
eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=1)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, 322, generator=g).cuda()
eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
lib = _lib.load()
out = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    eng.denoise(x, 500)
lib.mcm_debug_read(out, 1)
eng.denoise(x, 500)
lib.mcm_debug_read(out, 1)
names = ["wait accumulator", "wait staging free", "tmem ld (+bias stage)", "math", "stage + fence + issue", "chunks"]
tot = sum(out[i] for i in range(5))
print("one denoise step, 1 layer, all GEMMs: epilogue-warp cycles by phase")
for i, n in enumerate(names):
    if i < 5:
        print(f"  {n:24s} {out[i]/1e6:10.2f} Mcyc  {100*out[i]/tot:5.1f}%   per chunk {out[i]/max(1,out[5]):8.0f} cyc")
print("  chunks:", out[5])
