"""GPU: the fused cross-attention + FFN token kernel (csrc/fused_block.cu) -- phase by phase against float64
evaluations of the reference formulas, against the unfused kernel sequence, on ragged tile shapes, and end to end.

Tolerances (relative L2): an fp16 operand tile carries the 2^-11 rounding of its elements (~3-5e-4 relative L2 on
top of the upstream error), so a single phase is checked at 1e-3; block / sampler outputs at the 1e-3 bar of
BASELINE.json's north_star.
"""
import os

import numpy as np
import pytest
import torch

from motioncraft_b200 import synth
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
from tests import common as C

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _setup(T, B, layers=1):
    sd = C.hot(C.base_state(T, layers))
    g = torch.Generator().manual_seed(5)
    h0 = torch.randn(B, T, 512, generator=g)
    emb = torch.randn(B, 2048, generator=g)
    xf_out = torch.randn(B, 77, 256, generator=g)
    xf_proj = torch.randn(B, 2048, generator=g)
    eng = DenoiserEngine(sd, seq_len=T, num_layers=layers, max_batch=B)
    eng.set_option("fused_min_rows", 0)      # by default launches below 2048 rows take the kernel-per-op path
    eng.prepare_conditions(xf_out, xf_proj)
    return sd, h0, emb, xf_out, eng


@pytest.mark.parametrize("T,B", [(196, 4), (60, 5)])
def test_fused_phases_vs_float64(T, B):
    """Every phase product of the kernel (its shared-memory operand tile, h, the hidden scratch) against float64."""
    sd, h0, emb, xf_out, eng = _setup(T, B)
    want = C.block_stages(sd, h0, emb, xf_out)
    n = B * T * 512
    for k in (1, 2, 3, 4, 5, 6):
        eng.set_option("fused_stop", k)
        hk = eng.block_forward(0, 0, h0, emb)
        torch.cuda.synchronize()
        dump = eng.debug_copy(0, n).view(B, T, 512).float()
        if k in want:
            assert C.rel_l2(dump, want[k]) < TOL, k
        if k >= 4:
            assert C.rel_l2(hk, want[4]) < TOL, k
        if k == 5:    # B*T <= 74 * 256 rows: tile i runs on CTA pair i, so the scratch rows are in row order
            hid = eng.debug_copy(1, B * T * 1024).view(B, T, 1024).float()
            assert C.rel_l2(hid, want["hid"]) < TOL
    eng.set_option("fused_stop", 0)
    out = eng.block_forward(0, 0, h0, emb)
    assert C.rel_l2(out, want["out"]) < TOL
    assert torch.equal(out, eng.block_forward(0, 0, h0, emb)), "the fused kernel must be deterministic"
    eng.close()


@pytest.mark.parametrize("T,B", [(196, 1), (196, 7), (60, 1), (60, 9), (300, 3), (52, 6)])
def test_fused_matches_unfused_on_ragged_tiles(T, B):
    """Tile shapes the 256-row tiling has to survive: a single partial tile, a partial last tile, tiles spanning up to
    six samples (T=52), and a sample boundary in every lane quadrant.  Same arithmetic as the unfused sequence of
    GEMM / row kernels up to the fp32 summation order of the LayerNorm statistics."""
    sd, h0, emb, xf_out, eng = _setup(T, B)
    want = C.block_stages(sd, h0, emb, xf_out)["out"]
    eng.set_option("fused", 0)
    unfused = eng.block_forward(0, 0, h0, emb)
    eng.set_option("fused", 1)
    assert C.rel_l2(unfused, want) < TOL
    for sa_level in (0, 1, 2, 3):   # channel attention: separate kernels / fused tail / + fused head / + fused context
        eng.set_option("fused_sa", sa_level)
        fused = eng.block_forward(0, 0, h0, emb)
        assert C.rel_l2(fused, want) < TOL, sa_level
        assert C.rel_l2(fused, unfused) < 5e-4, sa_level
        assert torch.isfinite(fused).all()
    eng.close()


@pytest.mark.parametrize("fused", [0, 1])
def test_sampler_fused_and_unfused_vs_reference_golden(golden_dir, fused):
    """50-step DDIM at BASELINE config 0 shapes against the tensor produced by the unmodified reference, with the
    cross-attention + FFN either fused or as separate kernels (the round-1 path stays covered)."""
    g = np.load(os.path.join(golden_dir, "t2m_T60.npz"))
    x, xf_out, xf_proj = C.inputs(1, 60)
    eng = DenoiserEngine(C.hot(C.base_state(60)), seq_len=60, max_batch=1)
    eng.set_option("fused_min_rows", 0)
    eng.set_option("fused", fused)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    x0 = eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda())
    assert C.rel_l2(x0, g["ddim50_x0"]) < TOL
    eng.close()


def test_fused_benchmark_shape_batch_vs_oracle_sample():
    """BASELINE config 1 row count (B=256, T=196: 196 tiles over 74 CTA pairs, three waves): samples 0, 100 and 255 of the
    layer output against float64, the whole batch fused vs unfused."""
    T, B = 196, 256
    sd, h0, emb, xf_out, eng = _setup(T, B)
    fused = eng.block_forward(0, 0, h0, emb)
    eng.set_option("fused", 0)
    unfused = eng.block_forward(0, 0, h0, emb)
    assert C.rel_l2(fused, unfused) < 5e-4
    for b in (0, 100, 255):
        want = C.block_stages(sd, h0[b:b + 1], emb[b:b + 1], xf_out[b:b + 1])["out"]
        assert C.rel_l2(fused[b:b + 1], want) < TOL, b
    eng.close()


@pytest.mark.parametrize("T,B", [(196, 3), (256, 2), (64, 4), (40, 3), (300, 2)])
def test_channel_attention_fusion_levels(T, B):
    """sa_front_kernel / sa_tail_kernel against the kernel-per-op channel attention and float64, incl. the largest
    supported T (256: four full K blocks), heads of 10 / 16 / 49 / 64 features, and T = 300 where the fused kernels do not
    apply and the library must fall back to the unfused sequence by itself."""
    sd, h0, emb, xf_out, eng = _setup(T, B)
    want = C.block_stages(sd, h0, emb, xf_out)["out"]
    outs = {}
    for lvl in (0, 1, 2, 3):
        eng.set_option("fused_sa", lvl)
        outs[lvl] = eng.block_forward(0, 0, h0, emb)
        assert C.rel_l2(outs[lvl], want) < TOL, lvl
        assert torch.equal(outs[lvl], eng.block_forward(0, 0, h0, emb)), "deterministic"
    assert all(C.rel_l2(outs[lvl], outs[0]) < 5e-4 for lvl in (1, 2, 3))
    eng.close()


def test_control_branch_fused_vs_unfused_vs_oracle():
    """ControlT2MHalf_MCM: the copied blocks run through the same fused kernels on the control stream (a different h
    buffer and modulation slice per block); fused and kernel-per-op schedules against the fp32 oracle."""
    from motioncraft_b200 import modules
    T, B, n_ctrl, c_feats = 196, 3, 2, 35
    sd = synth.synth_state_dict(modules.ctrl_state_shapes(T, n_ctrl, c_feats))
    x, xf_out, xf_proj = C.inputs(B, T)
    c = synth.synth_tensor("c", (B, T, c_feats), synth.SEED_C_EMB)
    t = torch.full((B,), 640, dtype=torch.long)
    with torch.no_grad():
        want = O.control_forward(sd, x, t, xf_proj, xf_out, c)
    eng = DenoiserEngine(modules.engine_state_from_ctrl(sd), seq_len=T, max_batch=B, num_ctrl_blocks=n_ctrl,
                         ctrl_cond_feats=c_feats)
    eng.set_option("fused_min_rows", 0)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda(), c.cuda())
    outs = []
    for fused in (1, 0):
        eng.set_option("fused", fused)
        outs.append(eng.denoise(x.cuda(), 640))
        assert C.rel_l2(outs[-1], want) < TOL, fused
    assert C.rel_l2(outs[0], outs[1]) < 5e-4
    eng.close()


def test_small_launches_take_the_kernel_per_op_path():
    """Below `fused_min_rows` rows per launch the persistent tile kernels cannot fill the 74 CTA pairs and the library
    schedules the kernel-per-op sequence instead (B=1: 66 vs 75 ms per 50-step run); the switch must not change results
    beyond the fused / unfused difference and must follow the option."""
    T, B = 196, 2
    sd, h0, emb, xf_out, eng = _setup(T, B)
    fused = eng.block_forward(0, 0, h0, emb)                      # _setup set fused_min_rows = 0
    eng.set_option("fused_min_rows", 2048)
    small = eng.block_forward(0, 0, h0, emb)
    eng.set_option("fused", 0)
    unfused = eng.block_forward(0, 0, h0, emb)
    assert torch.equal(small, unfused), "392 rows < 2048: the default policy must pick the kernel-per-op path"
    assert not torch.equal(fused, unfused) and C.rel_l2(fused, unfused) < 5e-4
    eng.close()
