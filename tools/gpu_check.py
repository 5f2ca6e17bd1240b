"""Developer probe (GPU box): prints parity numbers of every stage against the CPU oracle.
Not a test and not a benchmark -- `pytest -m gpu` and bench.py are.  usage: python tools/gpu_check.py [T] [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncraft_b200 import modules, synth  # noqa: E402
from motioncraft_b200.engine import DenoiserEngine, SamplerTables, test_linear  # noqa: E402
from oracle import mcm_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def gemm_probe():
    g = torch.Generator().manual_seed(1)
    for (M, N, K) in [(128, 64, 64), (256, 256, 128), (300, 322, 512), (1000, 512, 322), (777, 196, 196),
                      (4096, 1024, 512), (130, 2440, 2048), (64, 40, 24)]:
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g)
        for fmt in (0, 1):
            C = test_linear(A.cuda(), W.cuda(), b.cuda(), fmt).cpu()
            if fmt == 0:
                ref = A.half().double() @ W.half().double().T + b.double()
            else:
                ref = A.double() @ W.double().T + b.double()
            print(f"gemm M={M} N={N} K={K} fmt={fmt}: rel={rel(C, ref):.3e} max|d|={(C.double() - ref).abs().max().item():.3e}",
                  flush=True)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    print(torch.cuda.get_device_name(0), flush=True)
    gemm_probe()
    shapes = modules.state_shapes(seq_len=T)
    sd = synth.synth_state_dict(shapes)
    sd = {k: v for k, v in sd.items() if ".ffn_channel." not in k}
    x = synth.synth_tensor("x_T", (B, T, 322), synth.SEED_XT)
    xf_out = synth.synth_tensor("xf_out", (B, 77, 256), synth.SEED_XF_OUT)
    xf_proj = synth.synth_tensor("xf_proj", (B, 2048), synth.SEED_XF_PROJ)
    sd64 = {k: v.double() for k, v in sd.items()}
    t = torch.full((B,), 999, dtype=torch.long)
    col = {}
    with torch.no_grad():
        e64 = O.mcm_forward(sd64, x.double(), t, xf_proj.double(), xf_out.double(), collect=col)
        e32 = O.mcm_forward(sd, x, t, xf_proj, xf_out)
    print(f"oracle fp32 vs fp64 eps rel = {rel(e32, e64):.3e}")
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    with torch.no_grad():
        x64 = O.ddim_sample_loop(lambda xx, tt: O.mcm_forward(sd64, xx, tt, xf_proj.double(), xf_out.double()),
                                 x.double(), tables, tmap)
    for precise in (True, False):
        eng = DenoiserEngine(sd, seq_len=T, max_batch=B, precise_all=precise)
        eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
        # block 0 alone
        hb = eng.block_forward(0, 0, col["h0"].float().cuda(), col["emb"].float().cuda())
        print(f"[precise={precise}] block0 rel = {rel(hb, col['h1']):.3e}", flush=True)
        eps = eng.denoise(x.cuda(), 999)
        print(f"[precise={precise}] eps(t=999) rel = {rel(eps, e64):.3e}", flush=True)
        eps2 = eng.denoise(x.cuda(), t.cuda())
        print(f"[precise={precise}] eps(t tensor) identical to uniform: {torch.equal(eps, eps2)}")
        st = SamplerTables(tables, tmap, "ddim")
        torch.cuda.synchronize()
        t0 = time.time()
        x0 = eng.sample(st, x.cuda())
        torch.cuda.synchronize()
        print(f"[precise={precise}] ddim50 x0 rel = {rel(x0, x64):.3e}  ({time.time() - t0:.3f}s)", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
