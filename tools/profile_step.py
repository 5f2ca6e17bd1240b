"""Profiling driver (GPU box): one warm denoise step of the bench workload between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [workload] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from motioncraft_b200 import modules, synth  # noqa: E402
from motioncraft_b200.engine import DenoiserEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "t2m"
    wl = dict(WORKLOADS[name])
    if len(sys.argv) > 2:
        wl["B"] = int(sys.argv[2])
    B, T = wl["B"], wl["T"]
    if wl["n_ctrl"]:
        sd = modules.engine_state_from_ctrl(synth.synth_state_dict(modules.ctrl_state_shapes(T, wl["n_ctrl"], wl["c_feats"])))
    else:
        sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
    eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_ctrl_blocks=wl["n_ctrl"], ctrl_cond_feats=wl["c_feats"])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, 322, generator=g).cuda()
    xf_out = torch.randn(B, 77, 256, generator=g).cuda()
    xf_proj = torch.randn(B, 2048, generator=g).cuda()
    c = torch.randn(B, wl["c_len"], wl["c_feats"], generator=g).cuda() if wl["n_ctrl"] else None
    eng.prepare_conditions(xf_out, xf_proj, c)
    for _ in range(2):
        eng.denoise(x, 500)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.denoise(x, 500)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    eng.close()


if __name__ == "__main__":
    main()
