"""Developer probe: time line of the GEMM launches of one small-batch denoise step (MCM_GEMM_TRACE, CUDA graph + PDL on).

    MCM_GEMM_TRACE=gpurun_out/gemm_trace.txt python tools/gemm_trace.py [B]       # writes the raw stamps at exit
    python tools/gemm_trace.py --read gpurun_out/gemm_trace.txt                  # per-launch phase table
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read(path):
    rows = []
    for ln in open(path):
        meta, st = ln.split("|")
        v = [int(x) for x in st.split()]
        rows.append((meta.strip(), v[:8], v[8:16], v[16:32], v[32:48], v[48:64]))
    rows = [r for r in rows if r[1][0] and r[1][7]]
    rows.sort(key=lambda r: r[1][0])
    print("per launch (ns, globaltimer): start->pdl_wait_done | ->first slab | ->mma issued | ->acc ready | ->stores issued | ->stores done | ->exit   || gap to next start, next pdl_wait_done")
    tot = {}
    for i, (meta, g, c, mk, pk, fine) in enumerate(rows):
        d = [g[1] - g[0], g[2] - g[1], g[3] - g[2], g[4] - g[3], g[5] - g[4], g[6] - g[5], g[7] - g[6]]
        nxt = rows[i + 1][1] if i + 1 < len(rows) else None
        gap = (nxt[0] - g[7], nxt[1] - g[7]) if nxt else (0, 0)
        print(f"{meta[:70]:70s} " + " ".join(f"{x:6d}" for x in d) + f"  | total {g[7] - g[0]:6d} || {gap[0]:7d} {gap[1]:7d}")
        if "--kb" in sys.argv:
            base = c[1]
            print("      producer issue (clk after pdl_wait): " + " ".join(str(x - base) for x in pk if x))
            print("      mma full-wait done               : " + " ".join(str(x - base) for x in mk if x))
            if fine[0]:
                print("      mma warp, kb = 2: wait %d | fence %d | 4 x MMA issue %d | commit %d | syncwarp %d" % (
                    mk[2] - fine[0], fine[1] - mk[2], fine[2] - fine[1], fine[3] - fine[2], fine[4] - fine[3]))
        for k, x in zip(("wait_pdl", "first_slab", "mma", "acc", "epi", "drain", "exit"), d):
            tot[k] = tot.get(k, 0) + x
        tot["gap_exit_to_next_ready"] = tot.get("gap_exit_to_next_ready", 0) + gap[1]
    n = len(rows)
    print(f"{n} launches; span {(rows[-1][1][7] - rows[0][1][0]) / 1e3:.1f} us")
    for k, x in tot.items():
        print(f"  {k:24s} {x / n:8.0f} ns per launch")


if len(sys.argv) > 2 and sys.argv[1] == "--read":
    read(sys.argv[2])
    sys.exit(0)

import torch
from motioncraft_b200 import modules, synth
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from motioncraft_b200.diffusion import build_diffusion
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = 196
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
d = build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small", respace="15,15,8,6,6"))
st = SamplerTables(d._tables(), d.timestep_map, "ddim", 0.0)
eng = DenoiserEngine(sd, seq_len=T, max_batch=B)
x = torch.randn(B, T, 322).cuda()
eng.prepare_conditions(torch.randn(B, 77, 256).cuda(), torch.randn(B, 2048).cuda())
for _ in range(3):
    eng.sample(st, x)
torch.cuda.synchronize()
eng.close()
