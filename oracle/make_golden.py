"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Runs in the build container only (needs /root/reference, imported under oracle/ref_shim.py).  Inputs
and weights are pure functions of seeds (motioncraft_b200/synth.py), so only OUTPUTS are stored.

    python oracle/make_golden.py            # writes tests/golden/{t2m_T60,ctrl_T60,schedule,repaint_T60,wav_encoder}.npz
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motioncraft_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v


@contextlib.contextmanager
def scripted_randn_like(noises):
    """Replace torch.randn_like by a scripted sequence (p_sample draws exactly one per step, :685)."""
    it = iter(noises)
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: next(it).to(x)
    try:
        yield
    finally:
        torch.randn_like = orig


def inputs(B, T):
    return (synth.synth_tensor("x_T", (B, T, 322), synth.SEED_XT),
            synth.synth_tensor("xf_out", (B, 77, 256), synth.SEED_XF_OUT),
            synth.synth_tensor("xf_proj", (B, 2048), synth.SEED_XF_PROJ))


def t2m(T=60, B=1):
    ref = ref_shim.build_reference_mcm(T=T)
    sd = synth.synth_state_dict({k: v.shape for k, v in ref.state_dict().items()})
    ref.load_state_dict(sd)
    x, xf_out, xf_proj = inputs(B, T)
    kw = dict(motion_mask=torch.ones(B, T), motion_length=torch.full((B,), T), xf_proj=xf_proj, xf_out=xf_out)
    out = {"keys": np.array(sorted(sd.keys()))}
    with torch.no_grad():
        for t in (999, 500, 0):
            out[f"eps_t{t}"] = ref(x, torch.full((B,), t, dtype=torch.long), **kw).numpy()
        ddim = ref_shim.build_reference_diffusion("15,15,8,6,6")
        out["ddim50_x0"] = ddim.ddim_sample_loop(ref, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                                 model_kwargs=dict(kw, y={}), eta=0).numpy()
        ddpm = ref_shim.build_reference_diffusion("10")
        noise = synth.synth_tensor("step_noise", (10, B, T, 322), synth.SEED_STEP_NOISE)
        with scripted_randn_like([noise[i] for i in reversed(range(10))]):
            out["ddpm10_x0"] = ddpm.p_sample_loop(ref, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                                  model_kwargs=dict(kw, y={})).numpy()
        out["ddpm10_timestep_map"] = np.array(ddpm.timestep_map)
    np.savez_compressed(os.path.join(GOLD, f"t2m_T{T}.npz"), **out)
    print("t2m", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items() if v.dtype != np.dtype("<U1") and v.dtype.kind == "f"})


def ctrl(T=60, B=1, n_ctrl=2, cond_feats=35, c_len=None):
    """ControlT2MHalf_MCM as tools/m2d_test.py:372-381 builds it (condition_pre_encode=False)."""
    from torch import nn
    ref_shim.install()
    from mogen.models.transformers.controlnet_mcm import ControlT2MHalf_MCM
    base = ref_shim.build_reference_mcm(T=T)
    for name in ("clip", "text_pre_proj", "textTransEncoder", "text_ln"):
        setattr(base, name, nn.Identity())
    model_cfg = dict(sa_block_cfg=dict(type="EfficientSelfAttention", latent_dim=T, num_heads=4, dropout=0, time_embed_dim=2048),
                     ca_block_cfg=dict(type="EfficientCrossAttention", latent_dim=512, text_latent_dim=256, num_heads=4,
                                       dropout=0, time_embed_dim=2048),
                     ffn_cfg=dict(latent_dim=512, ffn_dim=1024, dropout=0, time_embed_dim=2048))
    cfg = AttrDict(model=dict(model=model_cfg),
                   condition_encode_cfg=dict(dataset_name="finedance", condition_pre_encode=False, condition_cfg=True))
    net = ControlT2MHalf_MCM(base, copy_blocks_num=n_ctrl, control_cond_feats=cond_feats, cfg=cfg)
    net.eval()
    hot = {k: v.shape for k, v in net.state_dict().items()}
    sd = synth.synth_state_dict(hot)
    nn.Module.load_state_dict(net, sd)
    c_len = c_len or T - 3
    x, xf_out, xf_proj = inputs(B, T)
    c = synth.synth_tensor("c_m2d", (B, c_len, cond_feats), synth.SEED_C_M2D)
    kw = dict(motion_mask=torch.ones(B, T), motion_length=torch.full((B,), T), xf_proj=xf_proj, xf_out=xf_out)
    out = {"keys": np.array(sorted(sd.keys()))}
    with torch.no_grad():
        out["eps_t999"] = net(x, torch.full((B,), 999, dtype=torch.long), c=c, **kw).numpy()
        out["eps_t999_noc"] = net(x, torch.full((B,), 999, dtype=torch.long), c=None, **kw).numpy()
        ddim = ref_shim.build_reference_diffusion("15,15,8,6,6")
        out["ddim50_x0"] = ddim.ddim_sample_loop(net, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                                 model_kwargs=dict(kw, y={}, c=c), eta=0).numpy()
    out["c_len"] = np.array(c_len)
    np.savez_compressed(os.path.join(GOLD, f"ctrl_T{T}.npz"), **out)
    print("ctrl", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items() if v.dtype.kind == "f"})


def schedule():
    ref_shim.install()
    from mogen.models.utils import gaussian_diffusion as gd
    out = {}
    for tag, resp in (("ddim50", "15,15,8,6,6"), ("ddpm10", "10"), ("full", None)):
        d = ref_shim.build_reference_diffusion(resp)
        out[f"{tag}_timestep_map"] = np.array(getattr(d, "timestep_map", list(range(1000))))
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"{tag}_{k}"] = getattr(d, k)
    out["space_fast27"] = np.array(sorted(gd.space_timesteps(1000, "fast27")))
    out["space_ddim25"] = np.array(sorted(gd.space_timesteps(1000, "ddim25")))
    np.savez_compressed(os.path.join(GOLD, "schedule.npz"), **out)
    print("schedule ok")


def wav_encoder(out_dim=64, audio_in=2, n_samples=16000, B=2):
    """WavEncoder (mogen/models/utils/blocks.py:53-71) in eval mode on a seeded waveform; also the reference's parameter
    names of ConditionEncoder / the full-size encoder's frame count for the s2g window (159 900 samples -> 297)."""
    ref_shim.install()
    from mogen.models.utils.blocks import WavEncoder
    enc = WavEncoder(out_dim, audio_in=audio_in).eval()
    sd = synth.synth_state_dict({k: v.shape for k, v in enc.state_dict().items()})
    enc.load_state_dict(sd)
    wav = synth.synth_tensor("wav", (B, n_samples, audio_in), synth.SEED_C_EMB)
    with torch.no_grad():
        out = enc(wav).numpy()
        frames = WavEncoder(8, audio_in=audio_in).eval()(torch.zeros(1, 159900, audio_in)).shape[1]
    np.savez_compressed(os.path.join(GOLD, "wav_encoder.npz"), out=out, keys=np.array(sorted(sd.keys())),
                        frames_159900=np.array(frames), out_dim=np.array(out_dim), n_samples=np.array(n_samples))
    print("wav_encoder", out.shape, float(np.abs(out).max()), "frames(159900) =", frames)


def repaint(T=60, B=2, L=10):
    """RePaint / outpainting long-form sampling: SpacedDiffusion.ddim_sample_loop with
    y = {gt, outpainting_mask} as tools/m2d_test.py:176-195 builds it (first `overlap_len` frames kept), both through the
    harmonising loop (jump_length 3, jump_n_sample 5: 138 denoise + 108 undo steps) and with opt.no_repaint (plain loop,
    blend inside every ddim_sample).  torch.randn_like is scripted; the draw order is part of the contract."""
    import copy
    from oracle import mcm_oracle as O
    ref = ref_shim.build_reference_mcm(T=T)
    sd = synth.synth_state_dict({k: v.shape for k, v in ref.state_dict().items()})
    ref.load_state_dict(sd)
    x, xf_out, xf_proj = inputs(B, T)
    kw = dict(motion_mask=torch.ones(B, T), motion_length=torch.full((B,), T), xf_proj=xf_proj, xf_out=xf_out)
    gt = torch.zeros(T, 322)
    mask = torch.zeros(T, 322, dtype=torch.bool)
    gt[:L] = synth.synth_tensor("gt", (T, 322), synth.SEED_REPAINT_GT)[:L]
    mask[:L] = True
    out = {"overlap_len": np.array(L)}
    for mode in ("harmonize", "plain"):
        ddim = ref_shim.build_reference_diffusion("15,15,8,6,6")
        ddim.opt = copy.copy(ddim.opt)
        ddim.opt.overlap_len = L
        ddim.opt.no_repaint = (mode == "plain")
        times = None if mode == "plain" else O.schedule_jump_cjm_ddim(50, ddim.opt.jump_length, ddim.opt.jump_n_sample)
        n_den = 50 if times is None else sum(1 for a, b in zip(times[:-1], times[1:]) if b < a)
        n_undo = 0 if times is None else sum(1 for a, b in zip(times[:-1], times[1:]) if b >= a)
        n_draw = 2 * n_den + n_undo
        noise = synth.synth_tensor("repaint_noise", (n_draw, B, T, 322), synth.SEED_REPAINT_NOISE)
        with torch.no_grad(), scripted_randn_like([noise[i] for i in range(n_draw)]):
            out[f"{mode}_x0"] = ddim.ddim_sample_loop(ref, (B, T, 322), noise=x.clone(), clip_denoised=False,
                                                      model_kwargs=dict(kw, y={"gt": gt.clone(), "outpainting_mask": mask}),
                                                      eta=0).numpy()
        out[f"{mode}_n_draw"] = np.array(n_draw)
        if times is not None:
            out["times"] = np.array(times)
    np.savez_compressed(os.path.join(GOLD, f"repaint_T{T}.npz"), **out)
    print("repaint", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items() if v.dtype.kind == "f"})


TEXTS = ["a person walks forward and stumbles .", "jump", "the person is squatting , bending at the knees , keeping their "
         "back straight , while extending their arms ."]


def text_stack(B=3):
    """Trainable text-side stack (diffusion_transformer.py:157-171) of the reference with `clip_feat` supplied: the
    reference's own encode_text through the clip stand-in of ref_shim (frozen CLIP tower absent; EOT position from the
    stand-in tokenizer)."""
    ref = ref_shim.build_reference_mcm(T=60, num_layers=1, text_encoder=dict(ref_shim.TEXT_ENCODER_CFG))
    names = {k: v.shape for k, v in ref.state_dict().items()
             if k.startswith(("text_pre_proj.", "textTransEncoder.", "text_ln.", "text_proj."))}
    sd = synth.synth_state_dict(names)
    ref.load_state_dict(sd, strict=False)
    clip_feat = synth.synth_tensor("clip_feat", (B, 77, 512), synth.SEED_CLIP_FEAT)
    with torch.no_grad():
        xf_proj, xf_out = ref.encode_text(TEXTS[:B], clip_feat, "cpu")
        cond = ref.get_precompute_condition(text=TEXTS[:B], clip_feat=clip_feat, device="cpu")
    assert torch.equal(cond["xf_proj"], xf_proj) and torch.equal(cond["xf_out"], xf_out)
    eos = ref_shim.stub_clip_tokenize(TEXTS[:B]).argmax(dim=-1)
    np.savez_compressed(os.path.join(GOLD, "text_stack.npz"), xf_proj=xf_proj.numpy(), xf_out=xf_out.numpy(),
                        eos_index=eos.numpy(), keys=np.array(sorted(names.keys())), texts=np.array(TEXTS[:B]))
    print("text_stack: xf_proj", tuple(xf_proj.shape), "xf_out", tuple(xf_out.shape), "eos", eos.tolist())


def pathb(B=2, T=9, L=128):
    """Pieces of Path B the reference can pin here (SURVEY.md 8 rows b1, b3, b8, b9 + the start_x sampler): the reference's
    own PoseEncoder / PoseDecoder (motionx), the static human-topology mix of STMA, the CFG combine, and
    SpacedDiffusion(START_X, FIXED_LARGE) DDIM-50 / DDPM-10 around a fixed x_0-predictor."""
    ref_shim.install()
    import mogen.models.transformers.stmogen as st
    from mogen.models.utils import gaussian_diffusion as gd
    import torch.nn.functional as Fn
    enc = st.PoseEncoder(dataset_name="motionx", latent_dim=L, input_dim=322).eval()
    dec = st.PoseDecoder(dataset_name="motionx", latent_dim=L, output_dim=322).eval()
    names = {"joint_embed." + k: v.shape for k, v in enc.state_dict().items()}
    names.update({"out." + k: v.shape for k, v in dec.state_dict().items()})
    sd = synth.synth_state_dict(names)
    enc.load_state_dict({k[len("joint_embed."):]: v for k, v in sd.items() if k.startswith("joint_embed.")})
    dec.load_state_dict({k[len("out."):]: v for k, v in sd.items() if k.startswith("out.")})
    x = synth.synth_tensor("pb_motion", (B, T, 322), synth.SEED_XT)
    out = {"keys": np.array(sorted(names.keys()))}
    with torch.no_grad():
        h = enc(x)
        out["pose_encode"] = h.numpy()
        out["pose_decode"] = dec(h).numpy()
        bw = synth.synth_tensor("body_weight", (12, 12), synth.SEED_WEIGHTS)
        bv = h.reshape(B, T, 12, L)
        out["static_mix"] = torch.einsum("hl,bnld->bnhd", Fn.softmax(bw, dim=1), bv).numpy()       # st_attention.py:123-128
        a = synth.synth_tensor("cfg_text", (B, T, 322), synth.SEED_XT)
        b = synth.synth_tensor("cfg_none", (B, T, 322), synth.SEED_XF_OUT)
        holder = types.SimpleNamespace(scale_func_cfg=dict(scale=6.5))
        for t in (999, 500, 14, 0):
            coef = st.STMoGenTransformer.scale_func(holder, int(t))
            out[f"cfg_t{t}"] = (a * coef["text_coef"] + b * coef["none_coef"]).numpy()              # stmogen.py:755-759
        # start_x / fixed_large sampler around an x_0-predictor (the reference MCMTransformer used as a fixed function)
        Tm = 60
        net = ref_shim.build_reference_mcm(T=Tm, num_layers=2)
        msd = synth.synth_state_dict({k: v.shape for k, v in net.state_dict().items()})
        net.load_state_dict(msd)
        xT, xf_out, xf_proj = inputs(1, Tm)
        kw = dict(motion_mask=torch.ones(1, Tm), motion_length=torch.full((1,), Tm), xf_proj=xf_proj, xf_out=xf_out, y={})
        betas = gd.get_named_beta_schedule("linear", 1000)
        base = dict(betas=betas, model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_LARGE,
                    loss_type=gd.LossType.MSE)
        d50 = gd.SpacedDiffusion(use_timesteps=gd.space_timesteps(1000, "15,15,8,6,6"), opt=ref_shim.reference_opt(), **base)
        out["startx_ddim50_x0"] = d50.ddim_sample_loop(net, (1, Tm, 322), noise=xT, clip_denoised=False, model_kwargs=kw,
                                                       eta=0).numpy()
        d10 = gd.SpacedDiffusion(use_timesteps=gd.space_timesteps(1000, "10"), opt=ref_shim.reference_opt(), **base)
        noise = synth.synth_tensor("step_noise", (10, 1, Tm, 322), synth.SEED_STEP_NOISE)
        with scripted_randn_like([noise[i] for i in reversed(range(10))]):
            out["startx_ddpm10_x0"] = d10.p_sample_loop(net, (1, Tm, 322), noise=xT, clip_denoised=False,
                                                        model_kwargs=kw).numpy()
        # SFFN (stmogen.py:581-607) of the latent_dim = 64 configs (12 parts, D = 768), reduced ffn / time-embedding widths
        sf = st.SFFN(latent_dim=64, ffn_dim=128, dropout=0.0, time_embed_dim=256, num_heads=12).eval()
        sf_names = {"ffn." + k: v.shape for k, v in sf.state_dict().items()}
        sf_sd = synth.synth_state_dict(sf_names)           # (the zero-initialised output Linear gets real weights too)
        sf.load_state_dict({k[len("ffn."):]: v for k, v in sf_sd.items()})
        sx = synth.synth_tensor("sffn_x", (B, T, 768), synth.SEED_XT)
        se = synth.synth_tensor("sffn_emb", (B, 256), synth.SEED_XF_PROJ)
        out["sffn_keys"] = np.array(sorted(sf_names.keys()))
        out["sffn_out"] = sf(sx, se).numpy()
        # STMA.forward (st_attention.py:105-175) with its two mixture-of-experts layers replaced by stubs that return preset
        # tensors (tutel is un-vendored): everything AFTER the MoE outputs is the reference's own code.
        import mogen.models.attentions.st_attention as sta

        class _PresetMOE(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
                self.preset, self.aux_loss = None, 0.0

            def forward(self, x):
                return self.preset

        real_moe, sta.MOE = sta.MOE, _PresetMOE
        try:
            for tag, Ls, dyn, Ht, Tm in (("a", 64, True, 1, 20), ("b", 32, False, 12, 9)):
                m = sta.STMA(latent_dim=Ls, text_latent_dim=256, num_heads=12, num_text_heads=Ht, num_experts=16, topk=2,
                             gate_type="cosine_top", gate_noise=1.0, ffn_dim=4 * Ls, time_embed_dim=256, max_seq_len=196,
                             max_text_seq_len=77, temporal_comb=False, dropout=0.0, static_body=True, dynamic_body=dyn).eval()
                keep = {k: v for k, v in m.state_dict().items() if not k.startswith(("norm.", "text_norm."))}
                nm = {f"stma_{tag}." + k: v.shape for k, v in keep.items()}
                ssd = synth.synth_state_dict(nm)
                m.load_state_dict({k[len(f"stma_{tag}."):]: v for k, v in ssd.items()}, strict=False)
                Bs, Nt = 3, 7
                sx = synth.synth_tensor(f"stma_{tag}_x", (Bs, Tm, 12 * Ls), synth.SEED_XT)
                m.motion_moe.preset = synth.synth_tensor(f"stma_{tag}_mf", (Bs, Tm, 12, 4 * Ls), synth.SEED_XF_OUT)
                m.text_moe.preset = synth.synth_tensor(f"stma_{tag}_tf", (Bs, Nt, Ht, 2 * Ls), synth.SEED_C_EMB)
                semb = synth.synth_tensor(f"stma_{tag}_emb", (Bs, 256), synth.SEED_XF_PROJ)
                mask = torch.ones(Bs, Tm, 1)
                mask[1, Tm - 4:] = 0                                   # a padded sample
                cond = torch.tensor([1, 0, 11]).view(Bs, 1, 1)         # text on / off / on (cond_type % 10 > 0)
                xf = torch.zeros(Bs, Nt, Ht * 256)                     # only its shape is used once the MoE is preset
                out[f"stma_{tag}_keys"] = np.array(sorted(nm.keys()))
                out[f"stma_{tag}_out"] = m(sx, xf, semb, mask, cond, None, None).numpy()
        finally:
            sta.MOE = real_moe
    np.savez_compressed(os.path.join(GOLD, "pathb.npz"), **out)
    print("pathb", {k: v.shape for k, v in out.items() if v.dtype.kind == "f"})


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    schedule()
    t2m()
    ctrl()
    repaint()
    wav_encoder()
    text_stack()
    pathb()
