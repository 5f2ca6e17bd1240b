"""CPU: host-side logic -- registry, module/state_dict layout, schedule tables, the C-ABI library's
exported symbols, and the 'fail loudly without a GPU' contract."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import motioncraft_b200 as M
from motioncraft_b200 import _lib, diffusion, modules
from motioncraft_b200._lib import McmError
from oracle import mcm_oracle as O
from tests import common as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t2m_cfg(T=60, inference_type="ddim", respace="15,15,8,6,6"):
    dt = dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small")
    test = dict(dt, respace=respace) if respace else dict(dt)
    return dict(type="MotionDiffusion", model=dict(type="MCMTransformer", **modules.mcm_config(T)),
                loss_recon=dict(type="MSELoss", loss_weight=1, reduction="none"), diffusion_train=dt,
                diffusion_test=test, inference_type=inference_type)


def test_registry_resolves_config_types():
    for name in ("MotionDiffusion", "MCMTransformer", "EfficientSelfAttention", "EfficientCrossAttention", "MSELoss",
                 "ControlT2MHalf_MCM"):
        assert name in M.MODELS, name
    assert M.build_attention(None) is None
    with pytest.raises(KeyError):
        M.build_submodule(dict(type="NoSuchModel"))
    arch = M.build_architecture(_t2m_cfg())
    assert isinstance(arch, M.MotionDiffusion) and isinstance(arch.model, M.MCMTransformer)
    assert arch.diffusion_test.num_timesteps == 50
    arch2 = M.build_architecture(_t2m_cfg(inference_type="ddpm", respace=None))
    assert arch2.diffusion_test.num_timesteps == 1000


def test_state_dict_keys_and_zero_init(golden_dir):
    gold = np.load(os.path.join(golden_dir, "t2m_T60.npz"))
    m = M.MCMTransformer(**modules.mcm_config(60))
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(gold["keys"])           # identical to the reference's key set
    # the reference zero-initialises out, every out_layers[2] and every linear2
    assert float(sd["out.weight"].abs().max()) == 0.0
    assert float(sd["temporal_decoder_blocks.3.ffn_temporal.linear2.weight"].abs().max()) == 0.0
    assert float(sd["temporal_decoder_blocks.0.sa_block.proj_out.out_layers.2.weight"].abs().max()) == 0.0
    assert sd["temporal_decoder_blocks.0.sa_block.query.weight"].shape == (60, 60)
    assert sd["temporal_decoder_blocks.0.ca_block.key.weight"].shape == (512, 256)
    assert sd["temporal_decoder_blocks.0.sa_block.proj_out.emb_layers.1.weight"].shape == (120, 2048)


def test_control_wrapper_state_dict_layout(golden_dir):
    gold = np.load(os.path.join(golden_dir, "ctrl_T60.npz"))
    base = M.MCMTransformer(**modules.mcm_config(60))
    cfg = dict(model=dict(model=modules.mcm_config(60)),
               condition_encode_cfg=dict(dataset_name="finedance", condition_pre_encode=False, condition_cfg=True))
    net = M.ControlT2MHalf_MCM(base, copy_blocks_num=2, control_cond_feats=35, cfg=cfg)
    assert sorted(net.state_dict().keys()) == list(gold["keys"])
    assert sorted(C.ctrl_shapes(60, 2, 35).keys()) == list(gold["keys"])
    # base-only checkpoints load into base_model (controlnet_mcm.py:364-376)
    net.load_state_dict(base.state_dict())


def test_sa_latent_dim_must_equal_seq_len():
    cfg = modules.mcm_config(60)
    cfg["sa_block_cfg"]["latent_dim"] = 64
    with pytest.raises(McmError):
        M.MCMTransformer(**cfg)


def test_diffusion_tables_match_oracle_and_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small", respace="15,15,8,6,6"))
    assert d.timestep_map == list(g["ddim50_timestep_map"])
    for k in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
              "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped", "betas"):
        np.testing.assert_array_equal(getattr(d, k), g[f"ddim50_{k}"], err_msg=k)
    assert diffusion.space_timesteps(1000, "fast27") == set(O.space_timesteps(1000, "fast27"))
    assert diffusion.space_timesteps(1000, "ddim25") == set(O.space_timesteps(1000, "ddim25"))
    with pytest.raises(ValueError):
        diffusion.space_timesteps(1000, "ddim999")


def test_unsupported_parameterisations_fail_loudly():
    with pytest.raises(McmError):
        diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="previous_x",
                                       model_var_type="learned_range"))
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small", respace="10"))
    with pytest.raises(McmError):
        d.ddim_sample_loop(None, (1, 60, 322), clip_denoised=False,
                           model_kwargs={"y": {"outpainting_mask": torch.ones(1, 60, 322, dtype=torch.bool)}})
    with pytest.raises(McmError):
        d.training_losses()


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "mcm_b200.h")).read()
    declared = set(re.findall(r"\b(mcm_[a-z0-9_]+)\s*\(", header))
    declared -= {"mcm_ctx", "mcm_config", "mcm_sampler"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()                      # raises if the .so is absent: there is no fallback
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mcm_version()
    assert lib.mcm_kernel_launches() >= 0


def test_struct_layouts_match_header():
    header = open(os.path.join(ROOT, "include", "mcm_b200.h")).read()
    cfg_body = re.search(r"typedef struct mcm_config \{(.*?)\} mcm_config;", header, re.S).group(1)
    fields = re.findall(r"^\s*int\s+([a-z_]+);", cfg_body, re.M)
    assert fields == [f[0] for f in _lib.McmConfig._fields_]
    smp_body = re.search(r"typedef struct mcm_sampler \{(.*?)\} mcm_sampler;", header, re.S).group(1)
    fields = re.findall(r"^\s*(?:const\s+)?(?:int|float|unsigned long long)\*?\s+([a-z_0-9]+);", smp_body, re.M)
    assert fields[-1] == "model_mean_type"
    assert fields == [f[0] for f in _lib.McmSampler._fields_]


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    m = M.MCMTransformer(**modules.mcm_config(60)).eval()
    x, xf_out, xf_proj = C.inputs(1, 60)
    with pytest.raises(McmError):
        m(x, torch.zeros(1, dtype=torch.long), motion_mask=torch.ones(1, 60), xf_proj=xf_proj, xf_out=xf_out)
    lib = _lib.load()
    cfg = _lib.McmConfig(322, 60, 512, 2048, 1024, 256, 4, 8, 0, 0, 1, 77, 0)
    ctx = ctypes.c_void_p()
    assert lib.mcm_create(ctypes.byref(cfg), ctypes.byref(ctx)) != 0     # no device -> error status, not a crash
    assert len(lib.mcm_last_error()) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "motioncraft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "ref_shim" not in src and "/root/reference" not in src, f


def test_wav_encoder_matches_reference_golden(golden_dir):
    """Condition pre-encoder of the speech-to-gesture ControlNet branch: same parameter names as the
    reference's WavEncoder, bit-identical output on CPU for seeded weights / waveform (golden from the unmodified
    reference, oracle/make_golden.py::wav_encoder), and the s2g window length 159 900 samples -> 297 frames.  The encoder
    is torch / cuDNN library code evaluated once per sampling run, so it has a CPU evaluation to test against."""
    import numpy as np
    import os
    from motioncraft_b200 import synth
    from motioncraft_b200.condition_encoder import ConditionEncoder, WavEncoder
    g = np.load(os.path.join(golden_dir, "wav_encoder.npz"))
    enc = WavEncoder(int(g["out_dim"]), audio_in=2).eval()
    assert sorted(enc.state_dict().keys()) == list(g["keys"])
    enc.load_state_dict(synth.synth_state_dict({k: v.shape for k, v in enc.state_dict().items()}))
    wav = synth.synth_tensor("wav", (2, int(g["n_samples"]), 2), synth.SEED_C_EMB)
    with torch.no_grad():
        out = enc(wav)
    assert torch.equal(out, torch.from_numpy(g["out"]))
    assert WavEncoder(8, audio_in=2).eval()(torch.zeros(1, 159900, 2)).shape[1] == int(g["frames_159900"]) == 297
    ce = ConditionEncoder(dict(dataset_name="beats2", condition_pre_encode_type="wav", condition_latent_dim=32,
                               control_cond_feats=2))
    assert all(k.startswith("pre_encoder.feat_extractor.") for k in ce.state_dict())
    with pytest.raises(McmError):
        ConditionEncoder(dict(dataset_name="finedance", condition_pre_encode_type="wav", condition_latent_dim=32,
                              control_cond_feats=2))


def test_repaint_front_end_argument_checks():
    """Host logic of the outpainting branch of SpacedDiffusion.ddim_sample_loop: everything it needs from the tools' `opt`
    namespace, and every unsupported combination, is reported as McmError before any device work."""
    import argparse
    from motioncraft_b200 import scheduler
    cfg = dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small",
               respace="15,15,8,6,6")
    y = {"gt": torch.zeros(60, 322), "outpainting_mask": torch.ones(60, 322, dtype=torch.bool)}
    opt = argparse.Namespace(no_repaint=False, same_overlap_noisy=False, addBlend=True, overlap_len=10, no_resample=False,
                             timestep_respacing="ddim50", jump_length=3, jump_n_sample=5)
    with pytest.raises(McmError, match="opt"):                         # the reference dereferences self.opt (:962)
        diffusion.build_diffusion(cfg).ddim_sample_loop(None, (1, 60, 322), clip_denoised=False, model_kwargs={"y": y})
    d = diffusion.build_diffusion(cfg, opt=opt)
    with pytest.raises(McmError, match="eta"):
        d.ddim_sample_loop(None, (1, 60, 322), clip_denoised=False, model_kwargs={"y": y}, eta=0.5)
    with pytest.raises(McmError, match="gt"):
        d.ddim_sample_loop(None, (1, 60, 322), clip_denoised=False,
                           model_kwargs={"y": {"outpainting_mask": y["outpainting_mask"]}})
    opt1000 = argparse.Namespace(**{**vars(opt), "timestep_respacing": "ddim1000"})   # the tools' default: 600 > 50 steps
    with pytest.raises(McmError, match="schedule"):
        diffusion.build_diffusion(cfg, opt=opt1000).ddim_sample_loop(None, (1, 60, 322), clip_denoised=False,
                                                                     model_kwargs={"y": y})
    noisy = argparse.Namespace(**{**vars(opt), "same_overlap_noisy": True})
    with pytest.raises(McmError, match="same_overlap_noisy"):
        diffusion.build_diffusion(cfg, opt=noisy).ddim_sample_loop(None, (1, 60, 322), clip_denoised=False,
                                                                   model_kwargs={"y": y})
    # an all-False mask is the plain sampler (gaussian_diffusion.py:962: `True in outpainting_mask`): reaches the model bind
    with pytest.raises(AttributeError):
        d.ddim_sample_loop(None, (1, 60, 322), clip_denoised=False,
                           model_kwargs={"y": {"gt": y["gt"], "outpainting_mask": torch.zeros(60, 322, dtype=torch.bool)}})
    # schedule bookkeeping
    times = scheduler.get_schedule_jump_cjm_ddim(50, jump_length=3, jump_n_sample=5)
    n_den = sum(1 for a, b in zip(times[:-1], times[1:]) if b < a)
    assert (n_den, len(times) - 1 - n_den) == (138, 108) and scheduler.count_draws(times, 50) == 2 * 138 + 108
    assert scheduler.get_schedule_jump_cjm_ddim(25)[0] == 14 and scheduler.get_schedule_jump_cjm_ddim(50)[0] == 29
    assert scheduler.get_schedule_jump_cjm_ddim(50) == list(range(29, -2, -1))      # no resampling: straight walk down


def test_text_stack_module_matches_reference_golden(golden_dir):
    """MCMTransformer.encode_text / get_precompute_condition with clip_feat + eos_index (no `clip` package needed) against
    the reference's encode_text output; same torch library modules -> bit-identical on the CPU."""
    g = np.load(os.path.join(golden_dir, "text_stack.npz"))
    m = M.MCMTransformer(**modules.mcm_config(60, num_layers=1, text_encoder=dict(modules.TEXT_ENCODER_CFG))).eval()
    from motioncraft_b200 import synth
    m.load_state_dict(synth.synth_state_dict(modules.text_state_shapes()), strict=False)
    B = g["xf_proj"].shape[0]
    clip_feat = synth.synth_tensor("clip_feat", (B, 77, 512), synth.SEED_CLIP_FEAT)
    cond = m.get_precompute_condition(text=list(g["texts"]), clip_feat=clip_feat, eos_index=torch.from_numpy(g["eos_index"]),
                                      device="cpu")
    assert torch.equal(cond["xf_proj"], torch.from_numpy(g["xf_proj"]))
    assert torch.equal(cond["xf_out"], torch.from_numpy(g["xf_out"]))
    with pytest.raises(McmError):                      # neither the clip package nor the EOT position: loud failure
        m.get_precompute_condition(text=list(g["texts"]), clip_feat=clip_feat, device="cpu")
