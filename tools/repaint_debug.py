"""Developer probe (GPU box): localise RePaint sampler mismatches against the CPU oracle on truncated schedules."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import synth
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
from tests import common as C
T, B, L = 60, 2, 10
sd = C.base_state(T); x, xf_out, xf_proj = C.inputs(B, T)
gt = torch.zeros(T, 322); mask = torch.zeros(T, 322, dtype=torch.bool)
gt[:L] = synth.synth_tensor("gt", (T, 322), synth.SEED_REPAINT_GT)[:L]; mask[:L] = True
tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
eng = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B)
eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
st = SamplerTables(tables, tmap, "ddim")
noise = synth.synth_tensor("repaint_noise", (400, B, T, 322), synth.SEED_REPAINT_NOISE)
full = O.schedule_jump_cjm_ddim(50, 3, 5)
for name, times in [("one denoise", [29, 28]), ("two denoise", [29, 28, 27]), ("den-undo-den", [3, 2, 3, 2]),
                    ("first 12", full[:12]), ("first 40", full[:40]), ("full", full), ("plain", None)]:
    with torch.no_grad():
        want = O.ddim_repaint_loop(lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out), x.clone(), tables, tmap,
                                   tables["betas"], gt, mask, [noise[i] for i in range(400)], times=times, overlap_len=L)
    got = eng.sample_repaint(st, x.cuda(), gt, mask, noise.cuda(), times=times, betas=tables["betas"], overlap_len=L)
    print(f"{name:14s} rel = {C.rel_l2(got, want):.3e}", flush=True)
