"""GPU: the CUDA path (through the C-ABI) against the CPU oracle and the reference-generated golden
vectors.  Tolerances (relative L2, fp32):
    fast path  (fp16 operands, what ships)          : 1e-3  -- BASELINE.json north_star's bar
    precise_all (every GEMM bf16x2-split, 3 passes) : 5e-5  -- separates kernel bugs from rounding
"""
import os

import numpy as np
import pytest
import torch

import motioncraft_b200 as M
from motioncraft_b200 import modules, synth
from motioncraft_b200._lib import McmError
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
from tests import common as C

pytestmark = pytest.mark.gpu
TOL_FAST, TOL_PRECISE = 1e-3, 5e-5


def _engine(T, B, precise, num_layers=8):
    return DenoiserEngine(C.hot(C.base_state(T, num_layers)), seq_len=T, max_batch=B, precise_all=precise,
                          num_layers=num_layers)


@pytest.mark.parametrize("precise,tol", [(False, TOL_FAST), (True, TOL_PRECISE)])
def test_golden_reference_vectors_T60(golden_dir, precise, tol):
    """BASELINE config 0 shapes (B=1, T=60): eps at three timesteps, 50-step DDIM and 10-step DDPM against
    tensors produced by the UNMODIFIED reference."""
    g = np.load(os.path.join(golden_dir, "t2m_T60.npz"))
    x, xf_out, xf_proj = C.inputs(1, 60)
    eng = _engine(60, 1, precise)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    for t in (999, 500, 0):
        assert C.rel_l2(eng.denoise(x.cuda(), t), g[f"eps_t{t}"]) < tol, t
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    x0 = eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda())
    assert C.rel_l2(x0, g["ddim50_x0"]) < tol
    assert C.max_rel(x0, g["ddim50_x0"]) < 3 * tol
    tables, tmap = O.spaced_tables(1000, "10")
    noise = synth.synth_tensor("step_noise", (10, 1, 60, 322), synth.SEED_STEP_NOISE)
    x0 = eng.sample(SamplerTables(tables, tmap, "ddpm"), x.cuda(), noise.cuda())
    assert C.rel_l2(x0, g["ddpm10_x0"]) < tol
    eng.close()


@pytest.mark.parametrize("T,B", [(196, 3), (300, 2), (60, 5)])
def test_forward_and_blocks_vs_oracle(T, B):
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    col = {}
    want = C.oracle_forward(sd, x, 777, xf_proj, xf_out, torch.float64, collect=col)
    for precise, tol in ((True, TOL_PRECISE), (False, TOL_FAST)):
        eng = _engine(T, B, precise)
        eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
        assert C.rel_l2(eng.denoise(x.cuda(), 777), want) < tol
        # per-sample timesteps (the nn.Module call convention) give the same answer as the uniform fast path
        t = torch.full((B,), 777, dtype=torch.long)
        assert torch.equal(eng.denoise(x.cuda(), t.cuda()), eng.denoise(x.cuda(), 777))
        # every DecoderLayer alone, fed the oracle's own input for that layer
        for i in (0, 3, 7):
            got = eng.block_forward(0, i, col[f"h{i}"].float().cuda(), col["emb"].float().cuda())
            assert C.rel_l2(got, col[f"h{i + 1}"]) < tol, (precise, i)
        eng.close()


def test_mixed_timesteps_per_sample():
    T, B = 60, 4
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    t = torch.tensor([0, 14, 500, 999])
    want = C.oracle_forward(sd, x, t, xf_proj, xf_out, torch.float64)
    eng = _engine(T, B, False)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    assert C.rel_l2(eng.denoise(x.cuda(), t.cuda()), want) < TOL_FAST
    eng.close()


def test_ddim50_T196_vs_oracle():
    T, B = 196, 2
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    want = C.oracle_ddim(sd, x, xf_proj, xf_out)           # fp32 CPU, the reference's own arithmetic
    for precise, tol in ((False, TOL_FAST), (True, TOL_PRECISE)):
        eng = _engine(T, B, precise)
        eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
        tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
        st = SamplerTables(tables, tmap, "ddim")
        x0 = eng.sample(st, x.cuda())
        assert C.rel_l2(x0, want) < tol
        # host-buffer entry (what bench.py's e2e leg times) returns the same bits
        xh = x.clone().pin_memory()
        assert torch.equal(eng.sample_host(st, xh), x0.cpu())
        # deterministic: a second run is bit-identical
        assert torch.equal(eng.sample(st, x.cuda()), x0)
        eng.close()


def test_ddim_eta_nonzero_and_short_text():
    """eta != 0 consumes per-step noise (gaussian_diffusion.py:839-852); fewer than 77 text tokens."""
    T, B, N = 60, 2, 20
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T, n_tokens=N)
    tables, tmap = O.spaced_tables(1000, "10")
    noise = synth.synth_tensor("step_noise", (10, B, T, 322), synth.SEED_STEP_NOISE)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        want = O.ddim_sample_loop(lambda xx, tt: O.mcm_forward(sd64, xx, tt, xf_proj.double(), xf_out.double()),
                                  x.double(), tables, tmap, eta=0.5, step_noise=noise.double())
    eng = _engine(T, B, False)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    x0 = eng.sample(SamplerTables(tables, tmap, "ddim", eta=0.5), x.cuda(), noise.cuda())
    assert C.rel_l2(x0, want) < TOL_FAST
    # without explicit noise the library draws it on the device, keyed by the sampler seed: finite and reproducible
    a = eng.sample(SamplerTables(tables, tmap, "ddim", eta=0.5, seed=3), x.cuda(), None)
    assert torch.isfinite(a).all() and not torch.equal(a, x0)
    assert torch.equal(a, eng.sample(SamplerTables(tables, tmap, "ddim", eta=0.5, seed=3), x.cuda(), None))
    eng.close()


@pytest.mark.parametrize("precise,tol", [(False, TOL_FAST), (True, TOL_PRECISE)])
def test_control_branch_vs_reference_golden(golden_dir, precise, tol):
    g = np.load(os.path.join(golden_dir, "ctrl_T60.npz"))
    T = 60
    sd = synth.synth_state_dict(C.ctrl_shapes(T, 2, 35))
    x, xf_out, xf_proj = C.inputs(1, T)
    c = synth.synth_tensor("c_m2d", (1, int(g["c_len"]), 35), synth.SEED_C_M2D)
    eng = DenoiserEngine(C.engine_state_from_ctrl(sd), seq_len=T, max_batch=1, precise_all=precise,
                         num_ctrl_blocks=2, ctrl_cond_feats=35)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda(), c.cuda())
    assert C.rel_l2(eng.denoise(x.cuda(), 999), g["eps_t999"]) < tol
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    assert C.rel_l2(eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda()), g["ddim50_x0"]) < tol
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda(), None)       # c=None: plain base model
    assert C.rel_l2(eng.denoise(x.cuda(), 999), g["eps_t999_noc"]) < tol
    eng.close()


def test_batch_rows_are_independent():
    """Size-independent property: no op mixes samples, so a sample's result does not depend on its batch
    (what makes the multi-GPU batch sharding exact)."""
    T, B = 60, 6
    x, xf_out, xf_proj = C.inputs(B, T)
    eng = _engine(T, B, False)
    tables, tmap = O.spaced_tables(1000, "10")
    st = SamplerTables(tables, tmap, "ddim")
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    full = eng.sample(st, x.cuda())
    eng.prepare_conditions(xf_out[2:5].cuda(), xf_proj[2:5].cuda())
    part = eng.sample(st, x[2:5].cuda())
    assert torch.equal(full[2:5], part)
    eng.close()


def test_module_api_motion_diffusion_forward():
    """The reference-facing call: build_architecture(cfg) -> model(**data) -> list of per-sample dicts."""
    T, B = 60, 2
    dt = dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small")
    cfg = dict(type="MotionDiffusion", model=dict(type="MCMTransformer", **modules.mcm_config(T)),
               loss_recon=dict(type="MSELoss", loss_weight=1, reduction="none"), diffusion_train=dt,
               diffusion_test=dict(dt, respace="15,15,8,6,6"), inference_type="ddim")
    arch = M.build_architecture(cfg)
    arch.model.use_text_proj = True
    sd = C.base_state(T)
    arch.model.load_state_dict(sd)
    arch = arch.cuda().eval()
    x, xf_out, xf_proj = C.inputs(B, T)
    out = arch(motion=torch.zeros(B, T, 322).cuda(), motion_mask=torch.ones(B, T).cuda(),
               motion_length=torch.full((B,), T).cuda(), motion_metas=[{"text": "a"}, {"text": "b"}],
               xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), inference_kwargs={"noise": x.cuda()}, return_loss=False)
    assert isinstance(out, list) and len(out) == B
    assert set(out[0]) >= {"motion", "pred_motion", "motion_length", "motion_mask", "pred_motion_length",
                           "pred_motion_mask", "text"}
    assert out[0]["pred_motion"].device.type == "cpu" and out[0]["pred_motion"].shape == (T, 322)
    want = C.oracle_ddim(sd, x, xf_proj, xf_out)
    got = torch.stack([o["pred_motion"] for o in out])
    assert C.rel_l2(got, want) < TOL_FAST
    # nn.Module call convention of the denoiser itself + per-block call (what ControlT2MHalf_MCM relies on)
    t = torch.full((B,), 500, dtype=torch.long).cuda()
    eps = arch.model(x.cuda(), t, motion_mask=torch.ones(B, T).cuda(), xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    assert C.rel_l2(eps, C.oracle_forward(sd, x, 500, xf_proj, xf_out)) < TOL_FAST
    col = {}
    C.oracle_forward(sd, x, 500, xf_proj, xf_out, collect=col)
    h1 = arch.model.temporal_decoder_blocks[0](x=col["h0"].cuda(), xf=xf_out.cuda(), emb=col["emb"].cuda(),
                                               src_mask=None, motion_length=None, num_intervals=1)
    assert C.rel_l2(h1, col["h1"]) < TOL_FAST
    # reloading weights invalidates the packed copies
    sd2 = {k: v * 0.5 for k, v in sd.items()}
    arch.model.load_state_dict(sd2)
    eps2 = arch.model(x.cuda(), t, motion_mask=torch.ones(B, T).cuda(), xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    assert C.rel_l2(eps2, C.oracle_forward(sd2, x, 500, xf_proj, xf_out)) < TOL_FAST


def test_control_module_api():
    T, B = 60, 1
    base = M.MCMTransformer(**modules.mcm_config(T))
    base.use_text_proj = True
    cfg = dict(model=dict(model=modules.mcm_config(T)),
               condition_encode_cfg=dict(dataset_name="finedance", condition_pre_encode=False, condition_cfg=True))
    net = M.ControlT2MHalf_MCM(base, copy_blocks_num=2, control_cond_feats=35, cfg=cfg)
    sd = synth.synth_state_dict(C.ctrl_shapes(T, 2, 35))
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x, xf_out, xf_proj = C.inputs(B, T)
    c = synth.synth_tensor("c_m2d", (B, 57, 35), synth.SEED_C_M2D)
    t = torch.full((B,), 999, dtype=torch.long)
    got = net(x.cuda(), t.cuda(), motion_mask=torch.ones(B, T).cuda(), c=c.cuda(), xf_proj=xf_proj.cuda(),
              xf_out=xf_out.cuda())
    with torch.no_grad():
        want = O.control_forward(sd, x, t, xf_proj, xf_out, c)
    assert C.rel_l2(got, want) < TOL_FAST


def test_errors_are_reported_not_fatal():
    eng = _engine(60, 2, False)
    x, xf_out, xf_proj = C.inputs(3, 60)
    with pytest.raises(McmError):
        eng.denoise(x[:2].cuda(), 5)                       # conditions not prepared
    with pytest.raises(McmError):
        eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())   # batch 3 > max_batch 2
    eng.close()
    with pytest.raises(McmError):
        DenoiserEngine({}, seq_len=60, max_batch=1)        # missing parameters


@pytest.mark.parametrize("T,n_ctrl,c_feats,c_len", [(300, 2, 2048, 297), (1024, 4, 35, 1024)])
def test_control_at_benchmark_shapes(T, n_ctrl, c_feats, c_len):
    """BASELINE configs 2 (s2g: T=300, 2 control blocks, pre-encoded audio embedding [B, 297, 2048]) and 3 (m2d: T=1024,
    4 control blocks, music features [B, 1024, 35]) at B=1 against the fp32 oracle."""
    sd = synth.synth_state_dict(modules.ctrl_state_shapes(T, n_ctrl, c_feats))
    x, xf_out, xf_proj = C.inputs(1, T)
    c = synth.synth_tensor("c", (1, c_len, c_feats), synth.SEED_C_EMB)
    t = torch.full((1,), 640, dtype=torch.long)
    with torch.no_grad():
        want = O.control_forward(sd, x, t, xf_proj, xf_out, c)
    eng = DenoiserEngine(modules.engine_state_from_ctrl(sd), seq_len=T, max_batch=1, num_ctrl_blocks=n_ctrl,
                         ctrl_cond_feats=c_feats)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda(), c.cuda())
    assert C.rel_l2(eng.denoise(x.cuda(), 640), want) < TOL_FAST
    eng.close()


def test_full_batch_properties_t2m():
    """BASELINE config 1 at FULL size (B=256, T=196): size-independent checks -- finite output, bit-determinism,
    and rows of the big batch equal the same samples run alone (batch independence), plus oracle parity of 2 rows."""
    T, B = 196, 256
    sd = C.base_state(T)
    x = synth.synth_rows("x_T", (T, 322), synth.SEED_XT, 0, B)
    xf_out = synth.synth_rows("xf_out", (77, 256), synth.SEED_XF_OUT, 0, B)
    xf_proj = synth.synth_rows("xf_proj", (2048,), synth.SEED_XF_PROJ, 0, B)
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    st = SamplerTables(tables, tmap, "ddim")
    eng = _engine(T, B, False)
    # small launches would switch to the kernel-per-op schedule (same arithmetic, different fp32 summation order in the
    # LayerNorm statistics); pin the schedule so the 2-sample run below is comparable BIT FOR BIT with the big batch
    eng.set_option("fused_min_rows", 0)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    x0 = eng.sample(st, x.cuda())
    assert torch.isfinite(x0).all()
    assert torch.equal(eng.sample(st, x.cuda()), x0)
    sel = [3, 200]
    eng.prepare_conditions(xf_out[sel].cuda(), xf_proj[sel].cuda())
    assert torch.equal(eng.sample(st, x[sel].cuda()), x0[sel])
    want = C.oracle_ddim(sd, x[sel], xf_proj[sel], xf_out[sel])
    assert C.rel_l2(x0[sel], want) < TOL_FAST
    eng.close()


@pytest.mark.parametrize("env", [{"MCM_GRAPH": "0", "MCM_DUAL": "0"}, {"MCM_GRAPH": "0", "MCM_DUAL": "1"},
                                 {"MCM_GRAPH": "1", "MCM_DUAL": "0"}, {"MCM_CHUNK": "2"}, {"MCM_PAIR": "1"}])
def test_execution_modes_give_identical_results(env, tmp_path):
    """Eager vs CUDA-graph replay, one vs two streams (batch halves), sample chunking and cta_group::2 pair tiles are
    scheduling choices: the sampled x_0 must be BIT-identical to the default mode (graph + dual stream).  Modes are read
    from the environment at context creation, so each runs in a fresh process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
from tests import common as C
T, B = 60, 5
x, xf_out, xf_proj = C.inputs(B, T)
eng = DenoiserEngine(C.hot(C.base_state(T)), seq_len=T, max_batch=B)
eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
tables, tmap = O.spaced_tables(1000, "10")
x0 = eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda())
x0b = eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda())      # second run replays the cached graph
assert torch.equal(x0, x0b)
torch.save(x0.cpu(), sys.argv[1])
""" % root
    outs = []
    for e_extra in ({}, env):
        e = dict(os.environ)
        for k in ("MCM_GRAPH", "MCM_DUAL", "MCM_CHUNK", "MCM_PAIR"):
            e.pop(k, None)
        e.update(e_extra)
        f = str(tmp_path / ("x0_%d.pt" % len(outs)))
        r = subprocess.run([sys.executable, "-c", code, f], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(torch.load(f))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode", ["harmonize", "plain"])
def test_repaint_sampler_vs_reference_golden(golden_dir, mode):
    """RePaint / outpainting long-form DDIM (mcm_sample_repaint): first 10 frames pinned to a ground-truth tail, through the
    harmonising loop (138 denoise + 108 undo steps) and the plain 50-step loop, against the unmodified reference's x_0 with
    the same scripted noise draws.  Through the reference-facing front end (SpacedDiffusion.ddim_sample_loop + opt)."""
    import argparse
    from motioncraft_b200 import diffusion
    g = np.load(os.path.join(golden_dir, "repaint_T60.npz"))
    T, B, L = 60, 2, int(g["overlap_len"])
    x, xf_out, xf_proj = C.inputs(B, T)
    gt = torch.zeros(T, 322)
    mask = torch.zeros(T, 322, dtype=torch.bool)
    gt[:L] = synth.synth_tensor("gt", (T, 322), synth.SEED_REPAINT_GT)[:L]
    mask[:L] = True
    n_draw = int(g[f"{mode}_n_draw"])
    noise = synth.synth_tensor("repaint_noise", (n_draw, B, T, 322), synth.SEED_REPAINT_NOISE)
    opt = argparse.Namespace(no_repaint=(mode == "plain"), same_overlap_noisy=False, addBlend=True, overlap_len=L,
                             no_resample=False, timestep_respacing="ddim50", jump_length=3, jump_n_sample=5)
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small", respace="15,15,8,6,6"), opt=opt)
    net = M.MCMTransformer(**modules.mcm_config(T))
    net.use_text_proj = True            # as the reference model the golden was generated with (oracle/ref_shim.py)
    net.load_state_dict(C.base_state(T))
    net = net.cuda().eval()
    kw = dict(motion_mask=torch.ones(B, T).cuda(), motion_length=torch.full((B,), T).cuda(), xf_proj=xf_proj.cuda(),
              xf_out=xf_out.cuda(), y={"gt": gt.cuda(), "outpainting_mask": mask.cuda()})
    x0 = d.ddim_sample_loop(net, (B, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw, eta=0,
                            repaint_noise=noise.cuda())
    want = torch.from_numpy(g[f"{mode}_x0"])
    assert C.rel_l2(x0, want) < TOL_FAST
    # the pinned frames end exactly on the ground truth blend of the last step (mask == True there)
    assert torch.isfinite(x0).all()


def test_speech_control_with_raw_audio_condition():
    """speech-to-gesture as the reference's tools call it: ControlT2MHalf_MCM with condition_pre_encode=True receives RAW
    audio (B, samples, 2); the WavEncoder runs once (torch / cuDNN), its output enters control_cond_input inside the CUDA
    library.  Checked against the oracle fed with the CPU evaluation of the same encoder, and against handing the
    pre-encoded embedding directly."""
    from motioncraft_b200.condition_encoder import WavEncoder
    T, B, latent = 60, 2, 64
    base = M.MCMTransformer(**modules.mcm_config(T))
    base.use_text_proj = True
    cfg = dict(model=dict(model=modules.mcm_config(T)),
               condition_encode_cfg=dict(dataset_name="beats2", condition_pre_encode=True, condition_pre_encode_type="wav",
                                         condition_latent_dim=latent, control_cond_feats=2, condition_cfg=True))
    net = M.ControlT2MHalf_MCM(base, copy_blocks_num=2, control_cond_feats=2, cfg=cfg)
    assert any(k.startswith("condition_pre_encoder.pre_encoder.feat_extractor.0.conv1") for k in net.state_dict())
    sd = synth.synth_state_dict({k: v.shape for k, v in net.state_dict().items()})
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x, xf_out, xf_proj = C.inputs(B, T)
    wav = synth.synth_tensor("wav", (B, 16000, 2), synth.SEED_C_EMB)
    t = torch.full((B,), 999, dtype=torch.long)
    kw = dict(motion_mask=torch.ones(B, T).cuda(), xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    got = net(x.cuda(), t.cuda(), c=wav.cuda(), **kw)
    enc = WavEncoder(latent, audio_in=2).eval()
    enc.load_state_dict({k[len("condition_pre_encoder.pre_encoder."):]: v for k, v in sd.items()
                         if k.startswith("condition_pre_encoder.pre_encoder.")})
    with torch.no_grad():
        c_emb = enc(wav)                                              # (B, 30, latent) on the CPU
        osd = {k: v for k, v in sd.items() if not k.startswith("condition_pre_encoder.")}
        want = O.control_forward(osd, x, t, xf_proj, xf_out, c_emb)
    assert c_emb.shape == (B, 30, latent)
    assert C.rel_l2(got, want) < TOL_FAST
    got2 = net(x.cuda(), t.cuda(), c=c_emb.cuda(), **kw)              # pre-encoded embedding is accepted as well
    assert C.rel_l2(got2, want) < TOL_FAST
