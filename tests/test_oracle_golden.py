"""CPU: the oracle restatement against vectors produced by the UNMODIFIED reference (oracle/make_golden.py)
and against the schedule known-answers of SURVEY.md section 8(a2')."""
import os

import numpy as np
import pytest
import torch

from motioncraft_b200 import synth
from oracle import mcm_oracle as O
from tests import common as C


@pytest.fixture(scope="module")
def gold_t2m(golden_dir):
    return np.load(os.path.join(golden_dir, "t2m_T60.npz"))


def test_space_timesteps_known_answers():
    want = [0, 14, 28, 43, 57, 71, 85, 99, 114, 128, 142, 156, 171, 185, 199, 200, 214, 228, 243, 257, 271, 285, 299,
            314, 328, 342, 356, 371, 385, 399, 400, 428, 457, 485, 514, 542, 571, 599, 600, 640, 680, 719, 759, 799,
            800, 840, 880, 919, 959, 999]
    assert O.space_timesteps(1000, "15,15,8,6,6") == want
    ten = O.space_timesteps(1000, "10")
    assert ten[0] == 0 and ten[1] == 111 and ten[-2] == 888 and ten[-1] == 999 and len(ten) == 10
    with pytest.raises(ValueError):
        O.space_timesteps(10, "20")


def test_spaced_tables_known_answers():
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    ac = tables["alphas_cumprod"]
    assert len(tmap) == 50
    np.testing.assert_allclose(ac[[0, 1, 48, 49]], [0.9999, 0.9964144, 8.912634e-05, 4.035830e-05], rtol=2e-7)
    np.testing.assert_allclose(tables["sqrt_recip_alphas_cumprod"][49], 157.41046, rtol=1e-7)
    np.testing.assert_allclose(tables["sqrt_recipm1_alphas_cumprod"][49], 157.40728, rtol=1e-7)


def test_schedule_matches_reference_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    for tag, resp in (("ddim50", "15,15,8,6,6"), ("ddpm10", "10"), ("full", None)):
        tables, tmap = O.spaced_tables(1000, resp)
        assert list(g[f"{tag}_timestep_map"]) == tmap
        for k, v in tables.items():
            np.testing.assert_array_equal(v, g[f"{tag}_{k}"], err_msg=f"{tag}:{k}")   # float64, bit-exact
    assert list(g["space_fast27"]) == O.space_timesteps(1000, "fast27")
    assert list(g["space_ddim25"]) == O.space_timesteps(1000, "ddim25")


def test_state_dict_layout_matches_reference(gold_t2m):
    assert sorted(C.base_state(60).keys()) == list(gold_t2m["keys"])


@pytest.mark.parametrize("t", [999, 500, 0])
def test_forward_bit_exact_vs_reference(gold_t2m, t):
    x, xf_out, xf_proj = C.inputs(1, 60)
    got = C.oracle_forward(C.base_state(60), x, t, xf_proj, xf_out)
    # identical torch ops in identical order on the same host -> identical bits; allow 1e-6 relative
    # for a different CPU / MKL code path on another machine.
    assert C.rel_l2(got, gold_t2m[f"eps_t{t}"]) < 1e-6


def test_ddim50_vs_reference(gold_t2m):
    x, xf_out, xf_proj = C.inputs(1, 60)
    got = C.oracle_ddim(C.base_state(60), x, xf_proj, xf_out)
    assert C.rel_l2(got, gold_t2m["ddim50_x0"]) < 2e-6


def test_ddpm10_vs_reference(gold_t2m):
    x, xf_out, xf_proj = C.inputs(1, 60)
    noise = synth.synth_tensor("step_noise", (10, 1, 60, 322), synth.SEED_STEP_NOISE)
    got = C.oracle_ddpm(C.base_state(60), x, xf_proj, xf_out, noise)
    assert list(gold_t2m["ddpm10_timestep_map"]) == O.space_timesteps(1000, "10")
    assert C.rel_l2(got, gold_t2m["ddpm10_x0"]) < 2e-6


def test_control_forward_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ctrl_T60.npz"))
    sd = synth.synth_state_dict(C.ctrl_shapes(60, 2, 35))
    assert sorted(sd.keys()) == list(g["keys"])
    x, xf_out, xf_proj = C.inputs(1, 60)
    c = synth.synth_tensor("c_m2d", (1, int(g["c_len"]), 35), synth.SEED_C_M2D)
    t = torch.full((1,), 999, dtype=torch.long)
    with torch.no_grad():
        got = O.control_forward(sd, x, t, xf_proj, xf_out, c)
        got_noc = O.control_forward(sd, x, t, xf_proj, xf_out, None)
    assert C.rel_l2(got, g["eps_t999"]) < 1e-6
    assert C.rel_l2(got_noc, g["eps_t999_noc"]) < 1e-6


def test_hoisted_cross_attention_context_is_exact():
    """The step-invariant K/V context split (what the CUDA path caches per run) changes nothing."""
    sd = C.base_state(60)
    x, xf_out, xf_proj = C.inputs(2, 60)
    h, emb = O.embed(x, torch.full((2,), 7, dtype=torch.long), xf_proj, sd)
    pfx = "temporal_decoder_blocks.0.ca_block"
    with torch.no_grad():
        a = O.efficient_cross_attention(h, xf_out, emb, sd, pfx, 4)
        ctx = O.cross_attention_context(xf_out, sd, pfx, 4)
        b = O.efficient_cross_attention(h, None, emb, sd, pfx, 4, context=ctx)
    assert torch.equal(a, b)


def test_dead_ffn_channel_has_no_effect(gold_t2m):
    """mcm.py:33-34 computes ffn_channel and discards it: perturbing its weights must change nothing
    (this is why neither the oracle nor the CUDA path evaluates it)."""
    sd = dict(C.base_state(60))
    for k in list(sd):
        if ".ffn_channel." in k:
            sd[k] = sd[k] + 1.0
    x, xf_out, xf_proj = C.inputs(1, 60)
    got = C.oracle_forward(sd, x, 999, xf_proj, xf_out)
    assert C.rel_l2(got, gold_t2m["eps_t999"]) < 1e-6


def test_repaint_oracle_matches_reference_golden(golden_dir):
    """RePaint / outpainting long-form sampling: the oracle's restatement of ddim_sample's blend branch,
    the harmonising loop and `undo` against tensors produced by the UNMODIFIED reference (oracle/make_golden.py::repaint,
    scripted randn_like), and the product's own schedule restatement against the reference's."""
    from motioncraft_b200 import scheduler
    g = np.load(os.path.join(golden_dir, "repaint_T60.npz"))
    T, B, L = 60, 2, int(g["overlap_len"])
    times = [int(t) for t in g["times"]]
    assert O.schedule_jump_cjm_ddim(50, 3, 5) == times
    assert scheduler.get_schedule_jump_cjm_ddim(50, jump_length=3, jump_n_sample=5) == times
    assert scheduler.count_draws(times, 50) == int(g["harmonize_n_draw"]) and scheduler.count_draws(None, 50) == int(g["plain_n_draw"])
    assert times[0] == 29 and times[-1] == -1 and len(times) == 247
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    gt = torch.zeros(T, 322)
    mask = torch.zeros(T, 322, dtype=torch.bool)
    gt[:L] = synth.synth_tensor("gt", (T, 322), synth.SEED_REPAINT_GT)[:L]
    mask[:L] = True
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    mode = "plain"       # the harmonising loop (246 model calls) is exercised on the GPU; one CPU pass keeps this suite short
    n_draw = int(g[f"{mode}_n_draw"])
    noise = synth.synth_tensor("repaint_noise", (n_draw, B, T, 322), synth.SEED_REPAINT_NOISE)
    with torch.no_grad():
        got = O.ddim_repaint_loop(lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out), x.clone(), tables, tmap,
                                  tables["betas"], gt, mask, [noise[i] for i in range(n_draw)], times=None, overlap_len=L)
    assert torch.equal(got, torch.from_numpy(g[f"{mode}_x0"])), "oracle must be bit-identical to the reference on this host"


def test_text_stack_oracle_matches_reference_golden(golden_dir):
    """the trainable text-side stack (text_pre_proj -> 4-layer nn.TransformerEncoder -> text_ln
    -> text_proj at the EOT position, diffusion_transformer.py:157-171) restated in the oracle against outputs of the
    reference's own encode_text(clip_feat=...) -- bit for bit."""
    from motioncraft_b200 import modules
    g = np.load(os.path.join(golden_dir, "text_stack.npz"))
    shapes = modules.text_state_shapes()
    assert sorted(shapes.keys()) == list(g["keys"])
    sd = synth.synth_state_dict(shapes)
    B = g["xf_proj"].shape[0]
    clip_feat = synth.synth_tensor("clip_feat", (B, 77, 512), synth.SEED_CLIP_FEAT)
    with torch.no_grad():
        xf_proj, xf_out = O.encode_text_stack(sd, clip_feat, torch.from_numpy(g["eos_index"]))
    assert torch.equal(xf_proj, torch.from_numpy(g["xf_proj"]))
    assert torch.equal(xf_out, torch.from_numpy(g["xf_out"]))
