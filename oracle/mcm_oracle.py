"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's denoising hot path (Path A, configs/mcm/*).

This is the parity ORACLE for motioncraft_b200: a functional, dtype-generic (fp32 / fp64) torch-CPU
restatement of

  * the MCM denoiser forward        (mogen/models/transformers/{diffusion_transformer,mcm}.py)
  * its attention / AdaLN / FFN     (mogen/models/attentions/efficient_attention.py,
                                     mogen/models/utils/stylization_block.py)
  * the MCM ControlNet forward      (mogen/models/transformers/controlnet_mcm.py)
  * the DDIM / DDPM sampler tables and loops (mogen/models/utils/gaussian_diffusion.py)

Every function cites the reference lines it follows (paths relative to /root/reference).  It is
written with the SAME torch ops in the SAME order as the reference so that, in fp32 on the same
host, it reproduces the reference bit for bit; `oracle/validate_oracle.py` checks exactly that
against the unmodified reference imported under `oracle/ref_shim.py`, and `tests/golden/` holds
vectors generated from the reference itself (`oracle/make_golden.py`).

PINNING STATUS: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD
CONTAINER (tests/golden/*.npz, generator committed) plus the schedule known-answers of SURVEY.md
section 8(a2').

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (motioncraft_b200) never does; it fails loudly without its CUDA library.

`mm_round` (optional callable) is applied to both operands of every contraction; it exists to
EMULATE candidate tensor-core operand precisions (fp16 / tf32 / split-bf16) on the CPU before a
kernel is written (SURVEY.md section 7 "hard parts").
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _rnd(mm_round, x, tag):
    """Apply an operand-rounding emulation; a callable with `.wants_tag` also receives the name of
    the contraction (parameter prefix or einsum equation) so roundings can be applied selectively."""
    if getattr(mm_round, "wants_tag", False):
        return mm_round(x, tag)
    return mm_round(x)


def _lin(x, sd, prefix, mm_round=None):
    w, b = sd[prefix + ".weight"], sd.get(prefix + ".bias")
    if mm_round is not None:
        return F.linear(_rnd(mm_round, x, prefix), _rnd(mm_round, w, prefix), b)
    return F.linear(x, w, b)


def _ln(x, sd, prefix):
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[prefix + ".bias"], 1e-5)


def _einsum(eq, a, b, mm_round=None):
    if mm_round is not None:
        return torch.einsum(eq, _rnd(mm_round, a, eq), _rnd(mm_round, b, eq))
    return torch.einsum(eq, a, b)


def timestep_embedding(timesteps, dim, max_period=10000):
    """mogen/models/utils/position_encoding.py:42-60 (cos first, then sin; fp32 frequencies)."""
    half = dim // 2
    idx = torch.arange(start=0, end=half, dtype=torch.float32)
    freqs = torch.exp(-math.log(max_period) * idx / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
def stylization(h, emb, sd, prefix, mm_round=None):
    """StylizationBlock.forward, mogen/models/utils/stylization_block.py:29-40."""
    emb_out = _lin(F.silu(emb), sd, prefix + ".emb_layers.1", mm_round).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _ln(h, sd, prefix + ".norm") * (1 + scale) + shift
    return _lin(F.silu(h), sd, prefix + ".out_layers.2", mm_round)


def efficient_self_attention(x, emb, sd, prefix, num_heads, mm_round=None):
    """EfficientSelfAttention.forward, mogen/models/attentions/efficient_attention.py:25-46.

    Called by the MCM DecoderLayer on the TRANSPOSED tensor (tokens = 512 channels, features = T)
    with src_mask := ones (mcm.py:28-32), so the mask terms of :34 and :38 are exact no-ops
    (`+ (1-1)*-1e6` adds -0.0, `* 1`) and are omitted.
    """
    B, T, D = x.shape
    H = num_heads
    xn = _ln(x, sd, prefix + ".norm")
    query = _lin(xn, sd, prefix + ".query", mm_round)
    key = _lin(xn, sd, prefix + ".key", mm_round)
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    key = F.softmax(key.view(B, T, H, -1), dim=1)
    value = _lin(xn, sd, prefix + ".value", mm_round).view(B, T, H, -1)
    attention = _einsum("bnhd,bnhl->bhdl", key, value, mm_round)
    y = _einsum("bnhd,bhdl->bnhl", query, attention, mm_round).reshape(B, T, D)
    return x + stylization(y, emb, sd, prefix + ".proj_out", mm_round)


def cross_attention_context(xf, sd, prefix, num_heads, mm_round=None):
    """The step-invariant half of EfficientCrossAttention (efficient_attention.py:74-88,
    cond_type=None branch): depends only on the text features xf, not on x or t."""
    B, N, _ = xf.shape
    H = num_heads
    xfn = _ln(xf, sd, prefix + ".text_norm")
    key = _lin(xfn, sd, prefix + ".key", mm_round)
    key = F.softmax(key.view(B, N, H, -1), dim=1)
    value = _lin(xfn, sd, prefix + ".value", mm_round).view(B, N, H, -1)
    return _einsum("bnhd,bnhl->bhdl", key, value, mm_round)


def efficient_cross_attention(x, xf, emb, sd, prefix, num_heads, mm_round=None, context=None):
    """EfficientCrossAttention.forward, efficient_attention.py:64-92 (cond_type=None)."""
    B, T, D = x.shape
    H = num_heads
    query = _lin(_ln(x, sd, prefix + ".norm"), sd, prefix + ".query", mm_round)
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    attention = context if context is not None else cross_attention_context(xf, sd, prefix, H, mm_round)
    y = _einsum("bnhd,bhdl->bnhl", query, attention, mm_round).reshape(B, T, D)
    return x + stylization(y, emb, sd, prefix + ".proj_out", mm_round)


def ffn(x, emb, sd, prefix, mm_round=None):
    """FFN.forward, mogen/models/transformers/diffusion_transformer.py:25-28 (GELU = exact erf)."""
    y = _lin(F.gelu(_lin(x, sd, prefix + ".linear1", mm_round)), sd, prefix + ".linear2", mm_round)
    return x + stylization(y, emb, sd, prefix + ".proj_out", mm_round)


# The reference EXECUTES ffn_channel and throws its result away (mcm.py:33-34).  The parity oracle skips it (same values);
# the timed reference arm of bench.py sets this flag so that the CPU baseline does the work the reference really does.
EXECUTE_DEAD_FFN_CHANNEL = False


def decoder_layer(x, xf, emb, sd, prefix, num_heads, mm_round=None, ca_context=None):
    """mcm.py DecoderLayer.forward :25-41.  ffn_channel (:33-34) is computed and DISCARDED by the
    reference (its result never re-enters kwargs['x']), so it is not evaluated here unless the timing flag asks for it."""
    x = efficient_self_attention(x.transpose(-1, -2), emb, sd, prefix + ".sa_block", num_heads,
                                 mm_round).transpose(-1, -2)
    if EXECUTE_DEAD_FFN_CHANNEL and (prefix + ".ffn_channel.linear1.weight") in sd:
        ffn(x, emb, sd, prefix + ".ffn_channel", mm_round)                      # :33-34, result discarded as in the reference
    x = efficient_cross_attention(x, xf, emb, sd, prefix + ".ca_block", num_heads, mm_round, ca_context)
    return ffn(x, emb, sd, prefix + ".ffn_temporal", mm_round)


# ----------------------------------------------------------------------------------------------
# denoiser forward
# ----------------------------------------------------------------------------------------------
def _num_layers(sd, prefix="temporal_decoder_blocks."):
    n = 0
    while f"{prefix}{n}.sa_block.norm.weight" in sd:
        n += 1
    return n


def embed(motion, timesteps, xf_proj, sd, mm_round=None):
    """DiffusionTransformer.forward :206-218: emb = time_embed(sinusoid(t)) + xf_proj ; h = joint_embed(x)+pos."""
    D = sd["joint_embed.weight"].shape[0]
    T = motion.shape[1]
    te = timestep_embedding(timesteps, D).to(motion.dtype)
    emb = _lin(F.silu(_lin(te, sd, "time_embed.0", mm_round)), sd, "time_embed.2", mm_round)
    emb = emb + xf_proj
    h = _lin(motion, sd, "joint_embed", mm_round)
    h = h + sd["sequence_embedding"].unsqueeze(0)[:, :T, :]
    return h, emb


def mcm_forward(sd, motion, timesteps, xf_proj, xf_out, num_heads=4, mm_round=None, collect=None):
    """DiffusionTransformer.forward (diffusion_transformer.py:186-238) + MCMTransformer.forward_test
    (mcm.py:93-102), eval mode, use_text_proj=True, use_residual_connection=False."""
    B, T = motion.shape[:2]
    h, emb = embed(motion, timesteps, xf_proj, sd, mm_round)
    if collect is not None:
        collect["emb"] = emb
        collect["h0"] = h
    for i in range(_num_layers(sd)):
        h = decoder_layer(h, xf_out, emb, sd, f"temporal_decoder_blocks.{i}", num_heads, mm_round)
        if collect is not None:
            collect[f"h{i + 1}"] = h
    return _lin(h, sd, "out", mm_round).view(B, T, -1).contiguous()


def control_forward_c(sd, c, T, mm_round=None):
    """ControlT2MHalf_MCM.forward_c, controlnet_mcm.py:155-166, with `c` already pre-encoded
    (the WavEncoder output for s2g, the raw music features for m2d).  Keys use the wrapper's
    state_dict names: base_model.*, controlnet.{j}.*, control_cond_input.*"""
    c = _lin(c, sd, "control_cond_input", mm_round)
    len_c = c.shape[1]
    c_new = torch.cat([c, torch.zeros(c.shape[0], T - len_c, c.shape[2], dtype=c.dtype)], dim=-2)
    c_new[:, :len_c, :] = c_new[:, :len_c, :] + sd["base_model.sequence_embedding"].unsqueeze(0)[:, :len_c, :]
    return c_new


def control_forward(sd, motion, timesteps, xf_proj, xf_out, c, num_heads=4, mm_round=None):
    """ControlT2MHalf_MCM.forward + forward_test, controlnet_mcm.py:168-233, 306-361 (eval)."""
    base = {k[len("base_model."):]: v for k, v in sd.items() if k.startswith("base_model.")}
    B, T = motion.shape[:2]
    h, emb = embed(motion, timesteps, xf_proj, base, mm_round)
    n_total = _num_layers(base)
    n_ctrl = 0
    while f"controlnet.{n_ctrl}.after_proj.weight" in sd:
        n_ctrl += 1
    cc = control_forward_c(sd, c, T, mm_round) if c is not None else None
    h = decoder_layer(h, xf_out, emb, base, "temporal_decoder_blocks.0", num_heads, mm_round)
    start = 1
    if cc is not None:
        for index in range(1, n_ctrl + 1):
            j = index - 1
            pfx = f"controlnet.{j}"
            if j == 0:  # ControlT2MBlock.forward :65-75
                cin = h + _lin(cc, sd, pfx + ".before_proj", mm_round)
            else:       # :76-85
                cin = cc
            cc = decoder_layer(cin, xf_out, emb, sd, pfx + ".copied_block", num_heads, mm_round)
            c_skip = _lin(cc, sd, pfx + ".after_proj", mm_round)
            h = decoder_layer(h + c_skip, xf_out, emb, base, f"temporal_decoder_blocks.{index}",
                              num_heads, mm_round)
        start = n_ctrl + 1
    for index in range(start, n_total):
        h = decoder_layer(h, xf_out, emb, base, f"temporal_decoder_blocks.{index}", num_heads, mm_round)
    return _lin(h, base, "out", mm_round).view(B, T, -1).contiguous()


# ----------------------------------------------------------------------------------------------
# diffusion schedule (float64 numpy, once)
# ----------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------
# Trainable text-side stack (once per sampling run): DiffusionTransformer.encode_text with `clip_feat` supplied
# (mogen/models/transformers/diffusion_transformer.py:157-171; modules built at :123-145).  The frozen CLIP tower
# (:148-156) is an un-vendored third-party model (openai/CLIP, unpinned in requirements.txt:5) and is NOT restated:
# its output `clip_feat` (B, 77, 512) is an input here, and so is the EOT position `text.argmax(-1)` (:165).
# ---------------------------------------------------------------------------------------------
def text_encoder_layer(x, sd, prefix, nhead):
    """nn.TransformerEncoderLayer (post-norm, activation='gelu', dropout 0, batch_first=False: x is (N, B, E)) as
    torch.nn.modules.transformer evaluates it on its non-fused path (batch_first=False rules the fused path out)."""
    E = x.shape[-1]
    sa, _ = F.multi_head_attention_forward(
        x, x, x, E, nhead, sd[prefix + ".self_attn.in_proj_weight"], sd[prefix + ".self_attn.in_proj_bias"], None, None,
        False, 0.0, sd[prefix + ".self_attn.out_proj.weight"], sd[prefix + ".self_attn.out_proj.bias"], training=False,
        need_weights=False)
    x = F.layer_norm(x + sa, (E,), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"])
    ff = F.linear(F.gelu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"])),
                  sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])
    return F.layer_norm(x + ff, (E,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"])


def encode_text_stack(sd, clip_feat, eos_index, nhead=4):
    """clip_feat (B, 77, 512), eos_index (B,) -> (xf_proj (B, time_embed_dim), xf_out (B, 77, L))   [:157-171]"""
    x = clip_feat.permute(1, 0, 2)                                                        # :155
    if "text_pre_proj.weight" in sd:
        # x is a NON-contiguous permuted view here, so at::linear takes its matmul + add_ route, and at::matmul folds the
        # batch dimensions into one mm only when the weight `requires_grad` (it does in the reference: text_pre_proj is a
        # trainable nn.Linear) -- a different fp32 summation order from the batched route.  Reproduce that property.
        w = sd["text_pre_proj.weight"].detach().clone().requires_grad_(True)
        with torch.no_grad():
            x = F.linear(x, w, sd["text_pre_proj.bias"])                                 # :158
    n_layers = 1 + max([int(k.split(".")[2]) for k in sd if k.startswith("textTransEncoder.layers.")], default=-1)
    for i in range(n_layers):                                                             # :159
        x = text_encoder_layer(x, sd, f"textTransEncoder.layers.{i}", nhead)
    L = x.shape[-1]
    xf_out = F.layer_norm(x, (L,), sd["text_ln.weight"], sd["text_ln.bias"])             # :160
    sel = xf_out[eos_index.long(), torch.arange(xf_out.shape[1])]                         # :162-163
    xf_proj = F.linear(sel, sd["text_proj.0.weight"], sd["text_proj.0.bias"])
    return xf_proj, xf_out.permute(1, 0, 2)                                               # :165-166


def linear_beta_schedule(num_steps=1000):
    """get_named_beta_schedule('linear', n), gaussian_diffusion.py:235-253."""
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


def space_timesteps(num_timesteps, section_counts):
    """gaussian_diffusion.py:1346-1404 (returns a sorted list instead of a set)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return sorted(set(range(0, num_timesteps, i)))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        if section_counts == "fast27":
            steps = set(space_timesteps(num_timesteps, "15,15,8,6,6"))
            steps.remove(num_timesteps - 1)
            steps.add(num_timesteps - 3)
            return sorted(steps)
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac_stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            all_steps.append(start_idx + round(cur))
            cur += frac_stride
        start_idx += size
    return sorted(set(all_steps))


def diffusion_tables(betas):
    """GaussianDiffusion.__init__, gaussian_diffusion.py:354-387 (all float64)."""
    betas = np.array(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=post_var,
        posterior_log_variance_clipped=np.log(np.append(post_var[1], post_var[1:])),
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def spaced_tables(num_steps=1000, respace=None):
    """SpacedDiffusion.__init__, gaussian_diffusion.py:1416-1431: returns (tables, timestep_map)."""
    base = diffusion_tables(linear_beta_schedule(num_steps))
    if respace is None:
        return base, list(range(num_steps))
    use = set(space_timesteps(num_steps, respace))
    last = 1.0
    new_betas, tmap = [], []
    for i, a in enumerate(base["alphas_cumprod"]):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    return diffusion_tables(np.array(new_betas)), tmap


def _coef(arr, i, like):
    """_extract_into_tensor, gaussian_diffusion.py:1330-1343: float64 table -> indexed -> .float().
    (For an fp64 oracle run the value is kept in float64.)"""
    v = torch.from_numpy(np.asarray(arr))[i]
    return v.to(like.dtype) if like.dtype == torch.float64 else v.float()


# ----------------------------------------------------------------------------------------------
# sampler loops (epsilon prediction, fixed_small variance, clip_denoised=False)
# ----------------------------------------------------------------------------------------------
def ddim_sample_loop(model_fn, x_T, tables, tmap, eta=0.0, step_noise=None, trace=None, model_mean_type="epsilon"):
    """ddim_sample_loop_progressive + ddim_sample, gaussian_diffusion.py:799-852, 999-1049.

    `model_fn(x, t_original)` returns eps.  With eta = 0 (what MotionDiffusion passes,
    diffusion_architecture.py:184-191) sigma == 0, the per-step randn_like (:847) is multiplied by 0
    and the loop is deterministic given x_T.  No outpainting branch (y == {}).
    """
    x = x_T
    n = len(tmap)
    B = x.shape[0]
    for i in reversed(range(n)):
        t_model = torch.full((B,), tmap[i], dtype=torch.long)       # _WrappedModel, :1458-1463
        eps_model = model_fn(x, t_model)
        c1 = _coef(tables["sqrt_recip_alphas_cumprod"], i, x)
        c2 = _coef(tables["sqrt_recipm1_alphas_cumprod"], i, x)
        if model_mean_type == "start_x":                              # ModelMeanType.START_X :555-556 (configs/stmogen/*)
            pred_xstart = eps_model
        else:
            pred_xstart = c1 * x - c2 * eps_model                     # _predict_xstart_from_eps :572-577
        eps = (c1 * x - pred_xstart) / c2                             # _predict_eps_from_xstart :587-591
        alpha_bar = _coef(tables["alphas_cumprod"], i, x)
        alpha_bar_prev = _coef(tables["alphas_cumprod_prev"], i, x)
        sigma = (eta * torch.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) *
                 torch.sqrt(1 - alpha_bar / alpha_bar_prev))
        mean_pred = pred_xstart * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
        if eta != 0.0 and i != 0:
            mean_pred = mean_pred + sigma * step_noise[i]
        x = mean_pred
        if trace is not None:
            trace.append(x.clone())
    return x


def p_sample_loop(model_fn, x_T, tables, tmap, step_noise, trace=None, model_mean_type="epsilon", fixed_large=False):
    """p_sample_loop_progressive + p_sample + p_mean_variance (EPSILON / FIXED_SMALL),
    gaussian_diffusion.py:458-570, 634-696, 747-797.  `step_noise[i]` replaces the randn_like at
    :685 for retained step i (the reference draws from torch's global generator)."""
    x = x_T
    n = len(tmap)
    B = x.shape[0]
    for i in reversed(range(n)):
        t_model = torch.full((B,), tmap[i], dtype=torch.long)
        eps_model = model_fn(x, t_model)
        c1 = _coef(tables["sqrt_recip_alphas_cumprod"], i, x)
        c2 = _coef(tables["sqrt_recipm1_alphas_cumprod"], i, x)
        pred_xstart = eps_model if model_mean_type == "start_x" else c1 * x - c2 * eps_model
        mean = (_coef(tables["posterior_mean_coef1"], i, x) * pred_xstart +
                _coef(tables["posterior_mean_coef2"], i, x) * x)      # q_posterior_mean_variance :445-449
        if fixed_large:                                               # ModelVarType.FIXED_LARGE :527-531
            log_var = _coef(np.log(np.append(tables["posterior_variance"][1], tables["betas"][1:])), i, x)
        else:
            log_var = _coef(tables["posterior_log_variance_clipped"], i, x)
        nonzero = 0.0 if i == 0 else 1.0
        x = mean + nonzero * torch.exp(0.5 * log_var) * step_noise[i]
        if trace is not None:
            trace.append(x.clone())
    return x


# ----------------------------------------------------------------------------------------------
# RePaint / outpainting long-form sampling
# ----------------------------------------------------------------------------------------------
def schedule_jump_cjm_ddim(time_respacing=25, jump_length=1, jump_n_sample=1):
    """get_schedule_jump_cjm_ddim, mogen/models/utils/scheduler.py:178-208: the sequence of (respaced) timesteps the
    harmonising loop visits -- it starts at int(0.6 * time_respacing) - 1 (15 - 1 for 25), walks down to 0 and, at every
    `jump_length`-th step below t_T - jump_length, jumps back up `jump_length` steps `jump_n_sample - 1` times; ends with -1."""
    t_T = 15 if time_respacing == 25 else int(time_respacing * 0.6)
    jumps = {j: jump_n_sample - 1 for j in range(0, t_T - jump_length, jump_length)}
    t, ts = t_T, []
    while t >= 1:
        t -= 1
        ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] -= 1
            for _ in range(jump_length):
                t += 1
                ts.append(t)
    ts.append(-1)
    return ts


def repaint_step(x_in, eps_model, tables, i, gt, keep_mask, blend_noise, overlap_len, add_blend):
    """One ddim_sample call with eta = 0 and an outpainting mask, gaussian_diffusion.py:821-884 (same_overlap_noisy=False):
    the DDIM update, then x <- mask ? (sqrt(abar_prev) gt + sqrt(1 - abar_prev) noise) : x, with the overlap frames of the
    weighted ground truth linearly cross-faded into the sample once the noise weight drops below 0.2 (:872-876)."""
    x = x_in
    c1 = _coef(tables["sqrt_recip_alphas_cumprod"], i, x)
    c2 = _coef(tables["sqrt_recipm1_alphas_cumprod"], i, x)
    pred_xstart = c1 * x - c2 * eps_model
    eps = (c1 * x - pred_xstart) / c2
    alpha_bar_prev = _coef(tables["alphas_cumprod_prev"], i, x)
    sample = pred_xstart * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev) * eps      # sigma = 0
    x = sample
    if keep_mask is not None and bool(keep_mask.any()):
        noise_weight = torch.sqrt(1 - alpha_bar_prev)
        weighed_gt = torch.sqrt(alpha_bar_prev) * gt + noise_weight * blend_noise
        weighed_gt = weighed_gt.expand_as(x).clone()
        if bool(noise_weight < 0.2) and add_blend:
            lw = torch.linspace(0, 1, overlap_len).view(1, -1, 1).expand(x.shape[0], -1, -1).to(x.dtype)
            weighed_gt[:, :overlap_len, :] = weighed_gt[:, :overlap_len, :] * (1 - lw) + x[:, :overlap_len, :] * lw
        x = (weighed_gt * keep_mask) + (x * ~keep_mask)
    return x, pred_xstart


def ddim_repaint_loop(model_fn, x_T, tables, tmap, betas, gt, keep_mask, noise_seq, times=None, overlap_len=0,
                      add_blend=True):
    """ddim_sample_loop with y['outpainting_mask'] (gaussian_diffusion.py:925-997): the harmonising loop
    (:1050-1118) over `times` (denoise when the next time is lower, else `undo` :426-435 = re-noise with the respaced
    beta), or, with times=None (opt.no_repaint), the plain progressive loop -- ddim_sample blends in both.
    `noise_seq` replaces the reference's torch.randn_like draws IN ORDER: two per denoise call (the unused eta noise of
    :847, then the blend noise of :867 when the mask is set), one per undo."""
    x = x_T
    B = x.shape[0]
    it = iter(noise_seq)
    has_mask = keep_mask is not None and bool(keep_mask.any())

    def denoise(x, i):
        t_model = torch.full((B,), tmap[i], dtype=torch.long)
        eps_model = model_fn(x, t_model)
        next(it)                                              # :847 randn_like(x), multiplied by sigma = 0
        blend_noise = next(it) if has_mask else None          # :867
        return repaint_step(x, eps_model, tables, i, gt, keep_mask, blend_noise, overlap_len, add_blend)[0]

    if times is None:
        for i in reversed(range(len(tmap))):
            x = denoise(x, i)
        return x
    for t_last, t_cur in zip(times[:-1], times[1:]):
        if t_cur < t_last:
            x = denoise(x, t_last)
        else:
            beta = _coef(betas, t_last, x)
            x = torch.sqrt(1 - beta) * x + torch.sqrt(beta) * next(it)
    return x


# ----------------------------------------------------------------------------------------------
# candidate operand roundings for precision emulation
# ----------------------------------------------------------------------------------------------
def longform_windows(model_fn_for_window, n_windows, B, motion_length, pre_frames, overlap_len, tables, tmap, betas, mean, std,
                     x_T_list, noise_seq_list, times=None, add_blend=True, first_gt=None, feats=322):
    """The sliding-window driver of tools/m2d_test.py:145-222 / tools/s2g_test.py:144-241 with --repaint, for B sequences at
    once (the tools run one): window i is a complete sampling run whose first `overlap_len` frames are pinned, through
    y['gt'] / y['outpainting_mask'], to the tail of `outputs` -- the DE-NORMALISED prediction of window i - 1 (m2d_test.py:189,
    203-205) -- and the result is concat(pred_i[:round_l] for i < last, pred_last) (:207-210, :219-220).
    `model_fn_for_window(i)` returns the denoiser closure of window i (its text / control condition);
    mean / std are numpy arrays, `pred * std + mean` follows numpy's promotion (:203)."""
    import numpy as np
    round_l = motion_length - pre_frames
    outputs, out_motions = None, []
    for i in range(n_windows):
        gt = torch.zeros(B, motion_length, feats)
        keep = torch.zeros(B, motion_length, feats, dtype=torch.bool)
        if overlap_len > 0:
            if i == 0 and first_gt is not None:                                              # --fix_very_first (:183-186)
                keep[:, :overlap_len] = True
                gt[:, :overlap_len] = first_gt[:, :overlap_len]
            elif i > 0:                                                                       # :188-190
                keep[:, :overlap_len] = True
                gt[:, :overlap_len] = outputs[:, -overlap_len:]
        fn = model_fn_for_window(i)
        if bool(keep.any()):
            x0 = ddim_repaint_loop(fn, x_T_list[i], tables, tmap, betas, gt, keep, noise_seq_list[i], times=times,
                                   overlap_len=overlap_len, add_blend=add_blend)
        else:
            x0 = ddim_sample_loop(fn, x_T_list[i], tables, tmap)
        pred = x0.numpy() * std + mean                                                        # :203
        outputs = torch.tensor(pred)                                                          # :205
        out_motions.append(pred if i == n_windows - 1 else pred[:, :round_l])                # :207-210
    return np.concatenate(out_motions, axis=1)                                                # :219


def round_fp16(x):
    return x.half().to(x.dtype)


def round_bf16(x):
    return x.bfloat16().to(x.dtype)


def round_tf32_trunc(x):
    """tf32 as the tensor core reads an fp32 word: low 13 mantissa bits ignored."""
    return (x.float().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32).to(x.dtype)


def round_split_bf16x2(x):
    """hi + lo bf16 split (16 significant bits): what a 3-pass bf16 MMA effectively multiplies."""
    x32 = x.float()
    hi = x32.bfloat16().float()
    lo = (x32 - hi).bfloat16().float()
    return (hi + lo).to(x.dtype)
