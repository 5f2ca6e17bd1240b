"""Developer probe: cycles the fused kernel's compute warps and MMA warp spend per phase (MCM_FUSED_PROF=1)."""
import ctypes, os, sys
os.environ["MCM_FUSED_PROF"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import _lib, modules, synth
from motioncraft_b200.engine import DenoiserEngine
B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 196
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1)).items() if ".ffn_channel." not in k}
eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=1)
eng.set_option("dual", 0); eng.set_option("graph", 0); eng.set_option("fused_min_rows", 0)
g = torch.Generator().manual_seed(0)
h = torch.randn(B, T, 512, generator=g).cuda()
emb = torch.randn(B, 2048, generator=g).cuda()
eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
lib = _lib.load()
out = (ctypes.c_ulonglong * 32)()
for _ in range(2):
    eng.block_forward(0, 0, h, emb)
lib.mcm_debug_read32(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.block_forward(0, 0, h, emb); e1.record(); torch.cuda.synchronize()
lib.mcm_debug_read32(out, 1)
names = ["P0 LN->OPA", "wait G1", "E1 softmax", "wait G2 (rounds)", "E2 lnmod (rounds)", "wait G3", "E3 reduce", "E3 drain+bar",
         "P_CVT", "wait G4 (4 q)", "E4 gelu (4 q)", "E4 drain", "wait G5 (+stage)", "E5 lnmod", "wait G6", "E6 reduce"]
tiles = (B * T + 255) // 256
NCW = int(os.environ.get("MCM_FB_NCW", "8"))
nw = NCW * 2 * tiles     # compute warps x CTAs x tiles contributing
tot = sum(out[i] for i in range(16))
print(f"block_forward (incl. SA) {e0.elapsed_time(e1):.3f} ms; {tiles} pair tiles; per-warp-per-tile cycles by phase:")
for i, n in enumerate(names):
    print(f"  {n:22s} {out[i]/nw:10.0f} cyc  {100*out[i]/tot:5.1f}%")
print(f"  total per tile {(tot + out[19])/nw:10.0f} cyc")
first = min(tiles, 74)
print(f"  P0 of a CTA's FIRST tile: {out[19]/(NCW*2*first):.0f} cyc; later tiles: {out[0]/max(1,NCW*2*(tiles-first)):.0f} cyc")
print(f"MMA warp per tile: wait tempty {out[16]/tiles:.0f}, wait full {out[17]/tiles:.0f}, total {out[18]/tiles:.0f}")
print(f"E4 split per tile: TMEM load+wait {out[20]/nw:.0f}, bias+GELU {out[21]/nw:.0f}, staging wait+STS+fence+TMA issue {out[22]/nw:.0f}")
