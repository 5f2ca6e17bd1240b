"""Developer probe: per-phase cycles of sa_tail_kernel (MCM_FUSED_PROF=1 MCM_ST_PROF=1, fused_sa=2)."""
import ctypes, os, sys
os.environ["MCM_FUSED_PROF"] = "1"; os.environ["MCM_ST_PROF"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import _lib, modules, synth
from motioncraft_b200.engine import DenoiserEngine
B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 196
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1)).items() if ".ffn_channel." not in k}
eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=1)
eng.set_option("dual", 0); eng.set_option("graph", 0); eng.set_option("fused_sa", 2); eng.set_option("fused_min_rows", 0)
g = torch.Generator().manual_seed(0)
h = torch.randn(B, T, 512, generator=g).cuda(); emb = torch.randn(B, 2048, generator=g).cuda()
eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
lib = _lib.load(); out = (ctypes.c_ulonglong * 32)()
for _ in range(2): eng.block_forward(0, 0, h, emb)
lib.mcm_debug_read32(out, 1)
eng.block_forward(0, 0, h, emb)
lib.mcm_debug_read32(out, 1)
tiles = 2 * B
names = ["stage AdaLN params", "wait G_a", "E_a stats", "E_a normalise + arrive", "wait G_b", "E_b", "drain stores"]
for part in (0, 1):
    nw = 4 * 2 * tiles
    tot = sum(out[8 * part + i] for i in range(7))
    print(f"warps part {part}: per-warp-per-tile cycles (total {tot / nw:.0f})")
    for i, n in enumerate(names):
        print(f"   {n:20s} {out[8 * part + i] / nw:9.0f}")
