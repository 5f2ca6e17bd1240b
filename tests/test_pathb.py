"""First pieces of Path B (configs/stmogen/*, SURVEY.md section 8 rows b1, b3, b8, b9 + the start_x / fixed_large sampler):
the oracle against outputs of the UNMODIFIED reference (tests/golden/pathb.npz, oracle/make_golden.py::pathb) bit for bit on
the CPU, the CUDA path against the same goldens on the GPU.  The mixture-of-experts of row b2 (tutel) is parity-unpinned and
absent."""
import os

import numpy as np
import pytest
import torch

from motioncraft_b200 import synth
from oracle import mcm_oracle as O
from oracle import pathb_oracle as P
from tests import common as C

TOL_FAST, TOL_SPLIT = 1e-3, 5e-5       # fp16-operand denoiser / bf16x2-split GEMMs (embed / out class)


def _gold(golden_dir):
    return np.load(os.path.join(golden_dir, "pathb.npz"))


def _pose_sd(g):
    from motioncraft_b200 import pathb
    enc, dec = pathb.PoseEncoder(latent_dim=128), pathb.PoseDecoder(latent_dim=128)
    names = {"joint_embed." + k: tuple(v.shape) for k, v in enc.state_dict().items()}
    names.update({"out." + k: tuple(v.shape) for k, v in dec.state_dict().items()})
    assert sorted(names) == list(g["keys"])                    # the reference's parameter names and shapes
    return synth.synth_state_dict(names), enc, dec


def test_pathb_oracle_matches_reference_golden(golden_dir):
    g = _gold(golden_dir)
    sd, _, _ = _pose_sd(g)
    B, T = g["pose_encode"].shape[:2]
    x = synth.synth_tensor("pb_motion", (B, T, 322), synth.SEED_XT)
    with torch.no_grad():
        h = P.pose_encode(sd, x)
        assert torch.equal(h, torch.from_numpy(g["pose_encode"]))
        assert torch.equal(P.pose_decode(sd, h), torch.from_numpy(g["pose_decode"]))
        bw = synth.synth_tensor("body_weight", (12, 12), synth.SEED_WEIGHTS)
        assert torch.equal(P.static_body_mix(bw, h.reshape(B, T, 12, 128)), torch.from_numpy(g["static_mix"]))
        a = synth.synth_tensor("cfg_text", (B, T, 322), synth.SEED_XT)
        b = synth.synth_tensor("cfg_none", (B, T, 322), synth.SEED_XF_OUT)
        for t in (999, 500, 14, 0):
            assert torch.equal(P.cfg_combine(a, b, t, 6.5), torch.from_numpy(g[f"cfg_t{t}"])), t
    assert P.scale_func(999, 6.5)[0] == (1 - 1 / 1000) * 6.5 + 1 and sum(P.scale_func(123, 6.5)) == 1.0
    assert sorted(P.body_slice()) == list(range(322))           # the eleven part slices partition the 322 columns


def test_start_x_fixed_large_sampler_oracle_matches_reference_golden(golden_dir):
    g = _gold(golden_dir)
    T = 60
    sd = C.base_state(T, 2)
    x, xf_out, xf_proj = C.inputs(1, T)
    fn = lambda xx, tt: O.mcm_forward(sd, xx, tt, xf_proj, xf_out)  # noqa: E731  (used as a fixed x_0-predictor)
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    with torch.no_grad():
        got = O.ddim_sample_loop(fn, x, tables, tmap, model_mean_type="start_x")
        assert torch.equal(got, torch.from_numpy(g["startx_ddim50_x0"]))
        tables, tmap = O.spaced_tables(1000, "10")
        noise = synth.synth_tensor("step_noise", (10, 1, T, 322), synth.SEED_STEP_NOISE)
        got = O.p_sample_loop(fn, x, tables, tmap, noise, model_mean_type="start_x", fixed_large=True)
        assert torch.equal(got, torch.from_numpy(g["startx_ddpm10_x0"]))


def test_pathb_modules_fail_loudly_without_gpu_and_for_moe():
    from motioncraft_b200 import pathb
    from motioncraft_b200._lib import McmError
    with pytest.raises(McmError):
        pathb.STMoGenTransformer()
    # CPU tensors never reach a fallback: every Path-B module refuses them (the library is CUDA-only)
    with pytest.raises(McmError):
        pathb.PoseEncoder()(torch.zeros(1, 2, 322))
    with pytest.raises(McmError):
        pathb.SFFN(latent_dim=64, ffn_dim=128, dropout=0.0, time_embed_dim=256, num_heads=12)(torch.zeros(1, 2, 768), torch.zeros(1, 256))
    tail = pathb.STMATail(latent_dim=32, num_heads=12, num_text_heads=1, time_embed_dim=256, dynamic_body=True)
    with pytest.raises(McmError):
        tail(torch.zeros(1, 2, 384), torch.zeros(1, 2, 12, 128), torch.zeros(1, 3, 1, 64), torch.zeros(1, 256), torch.ones(1, 2, 1),
             torch.ones(1, 1, 1))
    # the reference's parameter names (st_attention.py:79-99 minus the MoE layers and the pre-MoE norms)
    assert {"body_weight", "body_d_attn.norm.weight", "body_d_attn.query.weight", "proj_out.emb_layers.1.weight",
            "proj_out.out_layers.2.bias"} <= set(tail.state_dict())


@pytest.mark.gpu
def test_gpu_pathb_pieces_vs_reference_golden(golden_dir):
    from motioncraft_b200 import pathb
    g = _gold(golden_dir)
    sd, enc, dec = _pose_sd(g)
    enc.load_state_dict({k[len("joint_embed."):]: v for k, v in sd.items() if k.startswith("joint_embed.")})
    dec.load_state_dict({k[len("out."):]: v for k, v in sd.items() if k.startswith("out.")})
    enc, dec = enc.cuda(), dec.cuda()
    B, T = g["pose_encode"].shape[:2]
    x = synth.synth_tensor("pb_motion", (B, T, 322), synth.SEED_XT)
    h = enc(x.cuda())
    assert h.shape == (B, T, 1536) and C.rel_l2(h, g["pose_encode"]) < TOL_SPLIT
    y = dec(torch.from_numpy(g["pose_encode"]).cuda())
    assert y.shape == (B, T, 322) and C.rel_l2(y, g["pose_decode"]) < TOL_SPLIT
    bw = synth.synth_tensor("body_weight", (12, 12), synth.SEED_WEIGHTS)
    mix = pathb.static_body_mix(bw.cuda(), torch.from_numpy(g["pose_encode"]).cuda().view(B, T, 12, 128))
    assert C.rel_l2(mix, g["static_mix"]) < 1e-6
    a = synth.synth_tensor("cfg_text", (B, T, 322), synth.SEED_XT)
    b = synth.synth_tensor("cfg_none", (B, T, 322), synth.SEED_XF_OUT)
    for t in (999, 500, 14, 0):
        assert torch.equal(pathb.cfg_combine(a.cuda(), b.cuda(), t).cpu(), torch.from_numpy(g[f"cfg_t{t}"])), t
    # a larger, ragged batch of rows through the block-structured GEMMs against the oracle
    x2 = synth.synth_tensor("pb_motion2", (3, 197, 322), synth.SEED_XT)
    with torch.no_grad():
        want = P.pose_decode(sd, P.pose_encode(sd, x2))
    assert C.rel_l2(dec(enc(x2.cuda())), want) < TOL_SPLIT


@pytest.mark.gpu
def test_gpu_start_x_fixed_large_sampler_vs_reference_golden(golden_dir):
    """ModelMeanType.START_X / ModelVarType.FIXED_LARGE (configs/stmogen/*): DDIM-50 and DDPM-10 through the front end
    (build_diffusion -> SpacedDiffusion -> mcm_sample) with the denoiser as a fixed x_0-predictor."""
    import motioncraft_b200 as M
    from motioncraft_b200 import diffusion, modules
    g = _gold(golden_dir)
    T = 60
    net = M.MCMTransformer(**modules.mcm_config(T, num_layers=2))
    net.use_text_proj = True
    net.load_state_dict(C.base_state(T, 2))
    net = net.cuda().eval()
    x, xf_out, xf_proj = C.inputs(1, T)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    cfg = dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="start_x", model_var_type="fixed_large")
    d = diffusion.build_diffusion(dict(cfg, respace="15,15,8,6,6"))
    x0 = d.ddim_sample_loop(net, (1, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw, eta=0)
    assert C.rel_l2(x0, g["startx_ddim50_x0"]) < TOL_FAST
    d = diffusion.build_diffusion(dict(cfg, respace="10"))
    noise = synth.synth_tensor("step_noise", (10, 1, T, 322), synth.SEED_STEP_NOISE)
    x0 = d.p_sample_loop(net, (1, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw, step_noise=noise.cuda())
    assert C.rel_l2(x0, g["startx_ddpm10_x0"]) < TOL_FAST


def _sffn_setup(g):
    from motioncraft_b200 import pathb
    mod = pathb.SFFN(latent_dim=64, ffn_dim=128, dropout=0.0, time_embed_dim=256, num_heads=12)
    names = {"ffn." + k: tuple(v.shape) for k, v in mod.state_dict().items()}
    assert sorted(names) == list(g["sffn_keys"])                # the reference's parameter names and shapes (stmogen.py:583-594)
    sd = synth.synth_state_dict(names)
    B, T = g["sffn_out"].shape[:2]
    x = synth.synth_tensor("sffn_x", (B, T, 768), synth.SEED_XT)
    emb = synth.synth_tensor("sffn_emb", (B, 256), synth.SEED_XF_PROJ)
    return mod, sd, x, emb


def test_sffn_oracle_matches_reference_golden(golden_dir):
    """SFFN + StylizationBlock (stmogen.py:581-607): the restatement against the unmodified reference module, bit for bit."""
    g = _gold(golden_dir)
    _, sd, x, emb = _sffn_setup(g)
    with torch.no_grad():
        got = P.sffn(sd, x, emb, 12, prefix="ffn.")
    assert torch.equal(got, torch.from_numpy(g["sffn_out"]))


@pytest.mark.gpu
def test_gpu_sffn_vs_reference_golden(golden_dir):
    """mcm_sffn_forward (two block-diagonal tcgen05 GEMM launches + AdaLN row kernel + output GEMM with the residual) against
    the reference SFFN's output; also at a batch that fills several tiles, against the oracle."""
    g = _gold(golden_dir)
    mod, sd, x, emb = _sffn_setup(g)
    mod.load_state_dict({k[len("ffn."):]: v for k, v in sd.items()})
    mod = mod.cuda()
    got = mod(x.cuda(), emb.cuda())
    want = torch.from_numpy(g["sffn_out"])
    assert C.rel_l2(got, want) < TOL_FAST, C.rel_l2(got, want)
    # the FFN branch alone (the residual x dominates the norm of the output): same tolerance on out - x
    assert C.rel_l2(got.cpu() - x, want - x) < 2 * TOL_FAST, C.rel_l2(got.cpu() - x, want - x)
    xb = synth.synth_tensor("sffn_xb", (5, 196, 768), synth.SEED_XT)
    eb = synth.synth_tensor("sffn_eb", (5, 256), synth.SEED_XF_PROJ)
    with torch.no_grad():
        wb = P.sffn(sd, xb, eb, 12, prefix="ffn.")
    gb = mod(xb.cuda(), eb.cuda())
    assert C.rel_l2(gb, wb) < TOL_FAST and C.rel_l2(gb.cpu() - xb, wb - xb) < 2 * TOL_FAST


_STMA_CASES = (("a", 64, True, 1, 20), ("b", 32, False, 12, 9))     # tag, latent_dim, dynamic_body, num_text_heads, T


def _stma_setup(g, tag, Ls, dyn, Ht, Tm):
    from motioncraft_b200 import pathb
    mod = pathb.STMATail(latent_dim=Ls, num_heads=12, num_text_heads=Ht, time_embed_dim=256, static_body=True, dynamic_body=dyn)
    names = {f"stma_{tag}." + k: tuple(v.shape) for k, v in mod.state_dict().items()}
    assert sorted(names) == list(g[f"stma_{tag}_keys"])        # the reference STMA's parameters minus its MoE / pre-MoE norms
    sd = synth.synth_state_dict(names)
    Bs, Nt = 3, 7
    x = synth.synth_tensor(f"stma_{tag}_x", (Bs, Tm, 12 * Ls), synth.SEED_XT)
    mf = synth.synth_tensor(f"stma_{tag}_mf", (Bs, Tm, 12, 4 * Ls), synth.SEED_XF_OUT)
    tf = synth.synth_tensor(f"stma_{tag}_tf", (Bs, Nt, Ht, 2 * Ls), synth.SEED_C_EMB)
    emb = synth.synth_tensor(f"stma_{tag}_emb", (Bs, 256), synth.SEED_XF_PROJ)
    mask = torch.ones(Bs, Tm, 1)
    mask[1, Tm - 4:] = 0
    cond = torch.tensor([1, 0, 11]).view(Bs, 1, 1)
    return mod, sd, (x, mf, tf, emb, mask, cond)


@pytest.mark.parametrize("tag,Ls,dyn,Ht,Tm", _STMA_CASES)
def test_stma_tail_oracle_matches_reference_golden(golden_dir, tag, Ls, dyn, Ht, Tm):
    """STMA.forward after its MoE layers (static + dynamic body, masked temporal linear attention over text + motion tokens,
    StylizationBlock): the restatement against the unmodified reference class run with preset MoE outputs, bit for bit."""
    g = _gold(golden_dir)
    _, sd, (x, mf, tf, emb, mask, cond) = _stma_setup(g, tag, Ls, dyn, Ht, Tm)
    with torch.no_grad():
        got = P.stma_tail(sd, x, mf, tf, emb, mask, cond, 12, Ls, True, dyn, prefix=f"stma_{tag}.")
    assert torch.equal(got, torch.from_numpy(g[f"stma_{tag}_out"]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,Ls,dyn,Ht,Tm", _STMA_CASES)
def test_gpu_stma_tail_vs_reference_golden(golden_dir, tag, Ls, dyn, Ht, Tm):
    g = _gold(golden_dir)
    mod, sd, (x, mf, tf, emb, mask, cond) = _stma_setup(g, tag, Ls, dyn, Ht, Tm)
    mod.load_state_dict({k[len(f"stma_{tag}."):]: v for k, v in sd.items()})
    mod = mod.cuda()
    got = mod(x.cuda(), mf.cuda(), tf.cuda(), emb.cuda(), mask.cuda(), cond.cuda()).cpu()
    want = torch.from_numpy(g[f"stma_{tag}_out"])
    assert torch.isfinite(got).all()
    assert C.rel_l2(got, want) < TOL_FAST, C.rel_l2(got, want)
    assert C.rel_l2(got - x, want - x) < 2 * TOL_FAST, C.rel_l2(got - x, want - x)      # the attention branch without the residual


@pytest.mark.gpu
def test_gpu_stma_tail_full_size_vs_oracle():
    """The latent_dim = 64 configuration at T = 196, 77 text tokens, B = 4 (several GEMM tiles per contraction), against the
    oracle; a text-off sample and a padded sample included."""
    from motioncraft_b200 import pathb
    Ls, Tm, Nt, Bs = 64, 196, 77, 4
    mod = pathb.STMATail(latent_dim=Ls, num_heads=12, num_text_heads=1, time_embed_dim=2048, static_body=True, dynamic_body=True)
    names = {"stma_f." + k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = synth.synth_state_dict(names)
    mod.load_state_dict({k[len("stma_f."):]: v for k, v in sd.items()})
    x = synth.synth_tensor("stma_f_x", (Bs, Tm, 12 * Ls), synth.SEED_XT)
    mf = synth.synth_tensor("stma_f_mf", (Bs, Tm, 12, 4 * Ls), synth.SEED_XF_OUT)
    tf = synth.synth_tensor("stma_f_tf", (Bs, Nt, 1, 2 * Ls), synth.SEED_C_EMB)
    emb = synth.synth_tensor("stma_f_emb", (Bs, 2048), synth.SEED_XF_PROJ)
    mask = torch.ones(Bs, Tm, 1)
    mask[2, 150:] = 0
    cond = torch.tensor([1, 1, 0, 21]).view(Bs, 1, 1)
    with torch.no_grad():
        want = P.stma_tail(sd, x, mf, tf, emb, mask, cond, 12, Ls, True, True, prefix="stma_f.")
    got = mod.cuda()(x.cuda(), mf.cuda(), tf.cuda(), emb.cuda(), mask.cuda(), cond.cuda()).cpu()
    assert C.rel_l2(got, want) < TOL_FAST and C.rel_l2(got - x, want - x) < 2 * TOL_FAST, (C.rel_l2(got, want), C.rel_l2(got - x, want - x))
