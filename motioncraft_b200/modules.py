"""nn.Module shells with the reference's registry names, constructor signatures and state_dict layout
(SURVEY.md section 8b), whose forward passes run inside the C-ABI CUDA library.

The modules hold parameters (so reference checkpoints load unchanged and `ControlT2MHalf_MCM`-style
attribute access keeps working) but contain NO arithmetic: `MCMTransformer.forward` and
`DecoderLayer.forward` hand device pointers to `libmcm_b200.so`.  There is no PyTorch fallback; a CPU
tensor or a missing library raises.

Reference classes mirrored (paths relative to the reference root):
  StylizationBlock          mogen/models/utils/stylization_block.py:14-40
  FFN                       mogen/models/transformers/diffusion_transformer.py:15-28
  EfficientSelfAttention    mogen/models/attentions/efficient_attention.py:9-46
  EfficientCrossAttention   mogen/models/attentions/efficient_attention.py:49-92
  DecoderLayer              mogen/models/transformers/mcm.py:12-41
  MCMTransformer            mogen/models/transformers/mcm.py:44-102 (+ DiffusionTransformer base)
  ControlT2MBlock / ControlT2MHalf_MCM   mogen/models/transformers/controlnet_mcm.py:29-87, 107-403
"""
import re

import numpy as np
import torch
from torch import nn

from ._lib import McmError
from .engine import DenoiserEngine
from .registry import ATTENTIONS, SUBMODULES, build_attention


def _zero_(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class _EngineOnly(nn.Module):
    """A parameter container whose arithmetic lives in the CUDA library."""

    def forward(self, *a, **k):  # pragma: no cover - never a compute path
        raise McmError(f"{type(self).__name__} has no stand-alone PyTorch forward in motioncraft_b200; it runs as "
                       "part of DecoderLayer / MCMTransformer inside libmcm_b200.so")


class StylizationBlock(_EngineOnly):
    def __init__(self, latent_dim, time_embed_dim, dropout):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * latent_dim))
        self.norm = nn.LayerNorm(latent_dim)
        self.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=dropout),
                                        _zero_(nn.Linear(latent_dim, latent_dim)))


class FFN(_EngineOnly):
    def __init__(self, latent_dim, ffn_dim, dropout, time_embed_dim):
        super().__init__()
        _check_dropout(dropout)
        self.linear1 = nn.Linear(latent_dim, ffn_dim)
        self.linear2 = _zero_(nn.Linear(ffn_dim, latent_dim))
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)


def _check_dropout(p):
    if p:
        raise McmError("motioncraft_b200 is an inference path: dropout must be 0 (configs/mcm/* set dropout = 0)")


@ATTENTIONS.register_module()
class EfficientSelfAttention(_EngineOnly):
    def __init__(self, latent_dim, num_heads, dropout, time_embed_dim=None):
        super().__init__()
        _check_dropout(dropout)
        if time_embed_dim is None:
            raise McmError("EfficientSelfAttention without time_embed_dim (the stmogen STMA use) is out of scope")
        self.num_heads = num_heads
        self.latent_dim = latent_dim
        self.norm = nn.LayerNorm(latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(latent_dim, latent_dim)
        self.value = nn.Linear(latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.time_embed_dim = time_embed_dim
        self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)


@ATTENTIONS.register_module()
class EfficientCrossAttention(_EngineOnly):
    def __init__(self, latent_dim, text_latent_dim, num_heads, dropout, time_embed_dim):
        super().__init__()
        _check_dropout(dropout)
        self.num_heads = num_heads
        self.latent_dim = latent_dim
        self.text_latent_dim = text_latent_dim
        self.norm = nn.LayerNorm(latent_dim)
        self.text_norm = nn.LayerNorm(text_latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(text_latent_dim, latent_dim)
        self.value = nn.Linear(text_latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = StylizationBlock(latent_dim, time_embed_dim, dropout)


class DecoderLayer(nn.Module):
    """mcm.py:12-41.  `ffn_channel` is kept as a parameter container (checkpoints carry its weights)
    but never evaluated: the reference computes it and discards the result (mcm.py:33-34)."""

    def __init__(self, sa_block_cfg=None, ca_block_cfg=None, ffn_cfg=None):
        super().__init__()
        if sa_block_cfg is None or ca_block_cfg is None or ffn_cfg is None:
            raise McmError("the MCM DecoderLayer needs sa_block_cfg, ca_block_cfg and ffn_cfg (configs/mcm/*)")
        self.sa_block = build_attention(sa_block_cfg)
        self.ca_block = build_attention(ca_block_cfg)
        self.ffn_channel = FFN(**ffn_cfg)
        self.ffn_temporal = FFN(**ffn_cfg)
        self._owner = None   # (weakref-free) set by the owning transformer: (owner, kind, index)

    def _bind(self, owner, kind, index):
        object.__setattr__(self, "_owner", (owner, kind, index))

    def forward(self, x=None, xf=None, emb=None, src_mask=None, **kwargs):
        if self._owner is None:
            raise McmError("DecoderLayer is not attached to an MCMTransformer / ControlT2MHalf_MCM engine")
        if kwargs.get("cond_type") is not None:
            raise McmError("cond_type (classifier-free training masks) is a training feature; out of scope")
        owner, kind, index = self._owner
        return owner._block_forward(kind, index, x, xf, emb)


def _shape_cfg(model):
    blk = model.temporal_decoder_blocks[0]
    return dict(seq_len=model.max_seq_len, input_feats=model.input_feats, latent_dim=model.latent_dim,
                time_embed_dim=model.time_embed_dim, ffn_dim=blk.ffn_temporal.linear1.out_features,
                text_latent_dim=blk.ca_block.text_latent_dim, num_heads=blk.ca_block.num_heads,
                num_layers=len(model.temporal_decoder_blocks))


@SUBMODULES.register_module()
class MCMTransformer(nn.Module):
    """DiffusionTransformer (diffusion_transformer.py:54-238) + MCMTransformer (mcm.py:44-102)."""

    def __init__(self, input_feats, max_seq_len=240, latent_dim=512, time_embed_dim=2048, num_layers=8,
                 sa_block_cfg=None, ca_block_cfg=None, ffn_cfg=None, text_encoder=None, use_pos_embedding=True,
                 use_residual_connection=False, time_embedding_type="sinusoidal", post_process_cfg=None,
                 init_cfg=None, max_batch=None, precise_all=False):
        super().__init__()
        self.init_cfg = init_cfg
        self.input_feats = input_feats
        self.max_seq_len = max_seq_len
        self.latent_dim = latent_dim
        self.num_layers = num_layers
        self.time_embed_dim = time_embed_dim
        self.use_pos_embedding = use_pos_embedding
        if not use_pos_embedding:
            raise McmError("use_pos_embedding=False is not used by configs/mcm/*; out of scope")
        if time_embedding_type != "sinusoidal":
            raise McmError("only the sinusoidal time embedding (configs/mcm/*) is implemented")
        if use_residual_connection:
            raise McmError("use_residual_connection=True is not used by configs/mcm/*; out of scope")
        if sa_block_cfg is not None and sa_block_cfg.get("latent_dim") != max_seq_len:
            raise McmError("MCM applies self-attention on the transposed tensor: sa_block_cfg.latent_dim must "
                           "equal max_seq_len (configs/mcm/mcm_t2m_smplx.py:44-45)")
        self.sequence_embedding = nn.Parameter(torch.randn(max_seq_len, latent_dim))
        self._build_text_encoder(text_encoder)
        self.joint_embed = nn.Linear(input_feats, latent_dim)
        self.time_embedding_type = time_embedding_type
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))
        self.temporal_decoder_blocks = nn.ModuleList(
            DecoderLayer(sa_block_cfg=sa_block_cfg, ca_block_cfg=ca_block_cfg, ffn_cfg=ffn_cfg)
            for _ in range(num_layers))
        self.out = _zero_(nn.Linear(latent_dim, input_feats))
        self.use_residual_connection = use_residual_connection
        self.post_process_cfg = post_process_cfg
        self._max_batch_hint = max_batch
        self._precise_all = precise_all
        self._engine = None
        self._engine_key = None
        self._param_epoch = 0          # bumped whenever the weights may have changed (load_state_dict, .to(), ...)
        self._packed_epoch = -1        # epoch the engine's packed copies were made from
        for i, blk in enumerate(self.temporal_decoder_blocks):
            blk._bind(self, 0, i)
        self._register_load_state_dict_pre_hook(self._params_changed_hook)

    # ------------------------------------------------------------------ text encoder (outside the hot path)
    def _build_text_encoder(self, text_encoder):
        # diffusion_transformer.py:109-145.  CLIP is not available offline; the modules that hold
        # trainable text-side weights are created so checkpoints load, CLIP itself is resolved lazily.
        self.use_text_proj = False
        self._text_cfg = text_encoder
        self.clip = None
        if text_encoder is None:
            return
        tl = text_encoder["latent_dim"]
        self.use_text_proj = text_encoder.get("use_text_proj", False)
        if text_encoder["pretrained_model"] != "clip":
            raise NotImplementedError(text_encoder["pretrained_model"])
        self.text_pre_proj = nn.Linear(512, tl) if tl != 512 else nn.Identity()
        n_layers = text_encoder.get("num_layers", 0)
        self.use_text_finetune = n_layers > 0
        if n_layers > 0:
            layer = nn.TransformerEncoderLayer(d_model=tl, nhead=text_encoder.get("num_heads", 4),
                                               dim_feedforward=text_encoder.get("ff_size", 2048),
                                               dropout=text_encoder.get("dropout", 0),
                                               activation=text_encoder.get("activation", "gelu"))
            self.textTransEncoder = nn.TransformerEncoder(layer, num_layers=n_layers)
        self.text_ln = nn.LayerNorm(tl)
        if self.use_text_proj:
            self.text_proj = nn.Sequential(nn.Linear(tl, self.time_embed_dim))

    def _clip_module(self):
        try:
            import clip
        except ImportError as e:
            raise McmError("tokenising text / running the frozen CLIP tower needs the `clip` package and its ViT-B/32 "
                           "weights, which are not available offline: pass precomputed xf_proj / xf_out (mcm.py:65), or "
                           "clip_feat together with eos_index (the position of CLIP's end-of-text token)") from e
        return clip

    def encode_text(self, text, clip_feat, device, eos_index=None):
        """diffusion_transformer.py:147-172 -- ONCE per sampling run, not part of the per-step hot path: the trainable
        stack text_pre_proj -> nn.TransformerEncoder -> text_ln -> text_proj runs through torch's library kernels on the
        model's device (the same policy as the WavEncoder).

        `eos_index` (B,) is an extension: the reference takes the end-of-text position from `clip.tokenize(text).argmax(-1)`
        (:165); callers that hold pre-computed `clip_feat` (the datasets' clip_feat_dir) can pass the position instead of
        having the `clip` package tokenise again."""
        tokens = None
        if eos_index is None or clip_feat is None:
            clip = self._clip_module()
            tokens = clip.tokenize(text, truncate=True).to(device)
            if eos_index is None:
                eos_index = tokens.argmax(dim=-1)
        dtype = self.text_ln.weight.dtype
        if clip_feat is None:
            if self.clip is None:
                self.clip, _ = clip.load("ViT-B/32", "cpu")
                for p in self.clip.parameters():
                    p.requires_grad = False
                self.clip = self.clip.to(device)
            with torch.no_grad():
                x = self.clip.token_embedding(tokens).type(self.clip.dtype)
                x = x + self.clip.positional_embedding.type(self.clip.dtype)
                x = self.clip.ln_final(self.clip.transformer(x.permute(1, 0, 2))).type(dtype)
        else:
            x = clip_feat.to(device=device, dtype=dtype).permute(1, 0, 2)
        with torch.no_grad():
            x = self.text_pre_proj(x)
            if self.use_text_finetune:
                x = self.textTransEncoder(x)
            xf_out = self.text_ln(x)
            eos_index = torch.as_tensor(eos_index, device=xf_out.device).long()
            if self.use_text_proj:
                xf_proj = self.text_proj(xf_out[eos_index, torch.arange(xf_out.shape[1], device=xf_out.device)])
                return xf_proj, xf_out.permute(1, 0, 2).contiguous()
            return None, xf_out.permute(1, 0, 2).contiguous()

    def get_precompute_condition(self, text=None, xf_proj=None, xf_out=None, device=None, clip_feat=None, eos_index=None,
                                 **kwargs):
        if xf_out is None or (xf_proj is None and self.use_text_proj):
            if self._text_cfg is None:
                raise McmError("this model was built without a text encoder: pass xf_proj / xf_out (mcm.py:65)")
            xf_proj, xf_out = self.encode_text(text, clip_feat, device, eos_index=eos_index)
        return {"xf_proj": xf_proj, "xf_out": xf_out}

    def post_process(self, motion):
        if self.post_process_cfg is not None:
            if self.post_process_cfg.get("unnormalized_infer", False):
                mean = torch.from_numpy(np.load(self.post_process_cfg["mean_path"])).type_as(motion)
                std = torch.from_numpy(np.load(self.post_process_cfg["std_path"])).type_as(motion)
            motion = motion * std + mean
        return motion

    # ------------------------------------------------------------------ engine plumbing
    def _params_changed_hook(self, *args, **kwargs):
        # the engine keeps packed 16-bit copies: they are re-packed IN PLACE (mcm_finalize_params again) at the next
        # use, the workspace / streams / graphs of the engine survive a load_state_dict
        self._param_epoch += 1

    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    def _hot_state_dict(self):
        skip = ("clip.", "text_pre_proj.", "textTransEncoder.", "text_ln.", "text_proj.")
        return {k: v for k, v in self.state_dict().items() if not k.startswith(skip) and ".ffn_channel." not in k}

    def _engine_extra(self):
        return {}, {}

    def engine(self, batch):
        dev = self.sequence_embedding.device
        if dev.type != "cuda":
            raise McmError("motioncraft_b200 modules compute on an sm_100a CUDA device only: move the model with "
                           ".cuda() (there is no CPU / PyTorch fallback path)")
        want = max(batch, self._max_batch_hint or 0)
        if self._engine is None or self._engine_key[0] != dev or self._engine.max_batch < batch:
            self._drop_engine()
            extra_sd, extra_cfg = self._engine_extra()
            sd = self._hot_state_dict()
            sd.update(extra_sd)
            self._engine = DenoiserEngine(sd, max_batch=want, precise_all=self._precise_all, device=dev,
                                          **_shape_cfg(self), **extra_cfg)
            self._engine_key = (dev,)
            self._packed_epoch = self._param_epoch
        elif self._packed_epoch != self._param_epoch:
            extra_sd, _ = self._engine_extra()
            sd = self._hot_state_dict()
            sd.update(extra_sd)
            self._engine.load_params(sd)
            self._packed_epoch = self._param_epoch
        return self._engine

    def _apply(self, fn, *a, **k):
        self._param_epoch += 1           # .to() / .cuda() / .float(): same device -> re-pack in place, new device -> new engine
        return super()._apply(fn, *a, **k)

    def _block_forward(self, kind, index, x, xf, emb):
        eng = self.engine(x.shape[0])
        if getattr(self, "_ft_zero", None) is None or self._ft_zero.shape[0] != x.shape[0] or self._ft_zero.device != x.device:
            self._ft_zero = torch.zeros(x.shape[0], self.time_embed_dim, device=x.device)
        eng.prepare_conditions_cached(xf, self._ft_zero)
        return eng.block_forward(kind, index, x, emb)

    # ------------------------------------------------------------------ forward
    def forward(self, motion, timesteps, motion_mask=None, motion_length=None, num_intervals=1, patch_size=1,
                **kwargs):
        """diffusion_transformer.py:186-238 (eval).  motion_mask / motion_length / num_intervals are accepted
        for signature parity; the MCM decoder layers replace the mask by ones (mcm.py:28) so it has no effect."""
        if self.training:
            raise McmError("motioncraft_b200 implements the inference path only: call .eval() first")
        if patch_size != 1:
            raise McmError("patch_size != 1 is not used by configs/mcm/*")
        cond = self.get_precompute_condition(device=motion.device, **kwargs)
        xf_proj = cond["xf_proj"] if self.use_text_proj else torch.zeros(
            motion.shape[0], self.time_embed_dim, device=motion.device)
        eng = self.engine(motion.shape[0])
        eng.prepare_conditions_cached(cond["xf_out"], xf_proj, self._control_condition(kwargs))
        return eng.denoise(motion, timesteps)

    def _control_condition(self, kwargs):
        return None

    def bind_for_sampling(self, batch, model_kwargs, device):
        """Used by GaussianDiffusion.{p,ddim}_sample_loop: resolve the conditions once, stage the
        step-invariant work in the engine, return the engine that runs the whole loop."""
        if self.training:
            raise McmError("motioncraft_b200 implements the inference path only: call .eval() first")
        cond = self.get_precompute_condition(device=device, **model_kwargs)
        xf_proj = cond["xf_proj"] if self.use_text_proj else torch.zeros(batch, self.time_embed_dim, device=device)
        eng = self.engine(batch)
        eng.prepare_conditions_cached(cond["xf_out"], xf_proj, self._control_condition(model_kwargs))
        return eng

    def forward_test(self, h=None, src_mask=None, emb=None, xf_out=None, **kwargs):
        """mcm.py:93-102: the decoder layers and `out` on an ALREADY EMBEDDED residual stream h (B, T, latent) with the
        time(+text) embedding emb (B, time_embed_dim).  src_mask is accepted for signature parity: DecoderLayer replaces
        it by ones (mcm.py:28), and the all-ones mask is an identity in every block.  One C call (mcm_layers_forward)."""
        if h is None or emb is None or xf_out is None:
            raise McmError("forward_test needs h, emb and xf_out (mcm.py:93-102)")
        eng = self.engine(h.shape[0])
        # xf_proj only enters through emb, which the caller already formed (diffusion_transformer.py:206-213)
        if getattr(self, "_ft_zero", None) is None or self._ft_zero.shape[0] != h.shape[0] or self._ft_zero.device != h.device:
            self._ft_zero = torch.zeros(h.shape[0], self.time_embed_dim, device=h.device)
        eng.prepare_conditions_cached(xf_out, self._ft_zero, self._control_condition(kwargs))
        return eng.layers_forward(h, emb)


def state_shapes(seq_len=196, input_feats=322, latent_dim=512, time_embed_dim=2048, ffn_dim=1024,
                 text_latent_dim=256, num_heads=4, num_layers=8):
    """name -> shape of the hot-path parameters of an MCMTransformer (text encoder excluded)."""
    with torch.device("meta"):
        m = MCMTransformer(**mcm_config(seq_len, input_feats, latent_dim, time_embed_dim, ffn_dim, text_latent_dim,
                                        num_heads, num_layers))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


TEXT_ENCODER_CFG = dict(pretrained_model="clip", latent_dim=256, num_layers=4, num_heads=4, ff_size=2048, dropout=0,
                        use_text_proj=True)      # configs/mcm/mcm_t2m_smplx.py:58-64


def text_state_shapes(seq_len=60, text_encoder=None):
    """name -> shape of the trainable text-side parameters (text_pre_proj / textTransEncoder / text_ln / text_proj)."""
    with torch.device("meta"):
        m = MCMTransformer(**mcm_config(seq_len, num_layers=1, text_encoder=dict(text_encoder or TEXT_ENCODER_CFG)))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()
            if k.startswith(("text_pre_proj.", "textTransEncoder.", "text_ln.", "text_proj."))}


def mcm_config(seq_len=196, input_feats=322, latent_dim=512, time_embed_dim=2048, ffn_dim=1024, text_latent_dim=256,
               num_heads=4, num_layers=8, text_encoder=None):
    """The `model=dict(...)` section of configs/mcm/mcm_t2m_smplx.py:37-66 as keyword arguments."""
    return dict(input_feats=input_feats, max_seq_len=seq_len, latent_dim=latent_dim, time_embed_dim=time_embed_dim,
                num_layers=num_layers,
                sa_block_cfg=dict(type="EfficientSelfAttention", latent_dim=seq_len, num_heads=num_heads, dropout=0,
                                  time_embed_dim=time_embed_dim),
                ca_block_cfg=dict(type="EfficientCrossAttention", latent_dim=latent_dim,
                                  text_latent_dim=text_latent_dim, num_heads=num_heads, dropout=0,
                                  time_embed_dim=time_embed_dim),
                ffn_cfg=dict(latent_dim=latent_dim, ffn_dim=ffn_dim, dropout=0, time_embed_dim=time_embed_dim),
                text_encoder=text_encoder)


# =================================================================================================
# ControlNet branch (speech-to-gesture / music-to-dance), controlnet_mcm.py
# =================================================================================================
def _cfg_get(cfg, *path):
    cur = cfg
    for key in path:
        cur = cur[key] if isinstance(cur, dict) else getattr(cur, key)
    return cur


class ControlT2MBlock(nn.Module):
    """controlnet_mcm.py:29-87: a trainable copy of base block `block_index` with zero-initialised
    before_proj (first block only) / after_proj."""

    def __init__(self, base_block, block_index=0, latent_dim=512, cfg=None):
        super().__init__()
        model_cfg = _cfg_get(cfg, "model", "model")
        self.copied_block = DecoderLayer(sa_block_cfg=_cfg_get(model_cfg, "sa_block_cfg"),
                                         ca_block_cfg=_cfg_get(model_cfg, "ca_block_cfg"),
                                         ffn_cfg=_cfg_get(model_cfg, "ffn_cfg"))
        self.copied_block.load_state_dict(base_block.state_dict())
        self.block_index = block_index
        self.hidden_size = latent_dim
        if block_index == 0:
            self.before_proj = _zero_(nn.Linear(latent_dim, latent_dim))
        self.after_proj = _zero_(nn.Linear(latent_dim, latent_dim))

    def forward(self, *a, **k):
        raise McmError("ControlT2MBlock runs inside ControlT2MHalf_MCM's fused forward in libmcm_b200.so")


class ControlT2MHalf_MCM(nn.Module):
    """controlnet_mcm.py:107-403: the base MCMTransformer plus `copy_blocks_num` control copies fed by a
    speech / music condition `c`.  The whole controlled forward (forward_c once per run, the interleaved
    base / control blocks per step) runs in the CUDA library."""

    def __init__(self, base_model, copy_blocks_num=2, control_cond_feats=438, cfg=None):
        super().__init__()
        self.cfg = cfg
        self.base_model = base_model.eval()
        self.copy_blocks_num = copy_blocks_num
        self.total_blocks_num = len(base_model.temporal_decoder_blocks)
        self.controlnet = nn.ModuleList(
            ControlT2MBlock(base_model.temporal_decoder_blocks[i], i, base_model.latent_dim, cfg)
            for i in range(copy_blocks_num))
        enc_cfg = _cfg_get(cfg, "condition_encode_cfg")
        self._pre_encode = bool(_cfg_get(enc_cfg, "condition_pre_encode"))
        if self._pre_encode:
            # controlnet_mcm.py:138-146: raw audio -> WavEncoder -> control_cond_input.  The encoder is step-invariant:
            # it runs ONCE per sampling run (cuDNN through torch, condition_encoder.py), not once per denoise step as in
            # the reference.  Callers may also hand the already encoded embedding as `c`.
            from .condition_encoder import ConditionEncoder
            in_feats = _cfg_get(enc_cfg, "condition_latent_dim")
            self.condition_pre_encoder = ConditionEncoder(enc_cfg)
        else:
            in_feats = control_cond_feats
            self.condition_pre_encoder = None
        self._enc_cache = None
        self.control_cond_feats = in_feats
        self.control_cond_input = _zero_(nn.Linear(in_feats, base_model.latent_dim))
        self._engine = None
        self._param_epoch = 0
        self._packed_epoch = (-1, -1)
        self._register_load_state_dict_pre_hook(self._params_changed_hook)
        self.eval()

    # expose what the architecture wrapper touches
    @property
    def time_embed_dim(self):
        return self.base_model.time_embed_dim

    @property
    def use_text_proj(self):
        return self.base_model.use_text_proj

    def get_precompute_condition(self, **kwargs):
        return self.base_model.get_precompute_condition(**kwargs)

    def post_process(self, output):
        return self.base_model.post_process(output)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def load_state_dict(self, state_dict, strict=True):
        # controlnet_mcm.py:364-376: a base-only checkpoint loads into base_model
        if all(k.startswith(("base_model", "controlnet", "control_cond_input", "condition_pre_encoder"))
               for k in state_dict.keys()):
            return super().load_state_dict(state_dict, strict)
        # the base model's own pre-hook bumps ITS epoch, which this module's engine also watches (engine())
        return self.base_model.load_state_dict(state_dict, strict)

    def _params_changed_hook(self, *args, **kwargs):
        self._param_epoch += 1
        self._enc_cache = None

    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._param_epoch += 1
        self._enc_cache = None
        return super()._apply(fn, *a, **k)

    def _engine_state(self):
        sd = self.base_model._hot_state_dict()
        for k, v in self.state_dict().items():
            if k.startswith(("controlnet.", "control_cond_input.")) and ".ffn_channel." not in k:
                sd[k] = v
        return sd

    def engine(self, batch):
        base = self.base_model
        dev = base.sequence_embedding.device
        if dev.type != "cuda":
            raise McmError("motioncraft_b200 modules compute on an sm_100a CUDA device only (no CPU fallback)")
        epoch = (self._param_epoch, base._param_epoch)      # own weights + the base model's (shared by this engine)
        if self._engine is None or self._engine.device != dev or self._engine.max_batch < batch:
            self._drop_engine()
            self._engine = DenoiserEngine(self._engine_state(), max_batch=max(batch, base._max_batch_hint or 0),
                                          precise_all=base._precise_all, device=dev,
                                          num_ctrl_blocks=self.copy_blocks_num,
                                          ctrl_cond_feats=self.control_cond_feats, **_shape_cfg(base))
            self._packed_epoch = epoch
        elif self._packed_epoch != epoch:
            self._engine.load_params(self._engine_state())
            self._packed_epoch = epoch
        return self._engine

    def _condition(self, c):
        if c is None:
            return None
        enc = self.condition_pre_encoder
        if enc is not None and c.shape[-1] == enc.raw_feats and c.shape[-1] != self.control_cond_feats:
            # identity cache: holds a strong reference to the raw audio tensor, so its address cannot be recycled for a
            # different batch while the entry is alive
            key = (id(c), c.data_ptr(), c._version, tuple(c.shape))
            if self._enc_cache is None or self._enc_cache[0] != key:
                dev = self.control_cond_input.weight.device
                self._enc_cache = (key, enc(c.to(device=dev, dtype=torch.float32)), c)
            c = self._enc_cache[1]
        if c.shape[-1] != self.control_cond_feats:
            raise McmError(f"control condition has {c.shape[-1]} features, control_cond_input expects "
                           f"{self.control_cond_feats} (raw audio must be pre-encoded; see INTEGRATION.md)")
        return c

    def bind_for_sampling(self, batch, model_kwargs, device):
        base = self.base_model
        cond = base.get_precompute_condition(device=device, **model_kwargs)
        xf_proj = cond["xf_proj"] if base.use_text_proj else torch.zeros(batch, base.time_embed_dim, device=device)
        eng = self.engine(batch)
        eng.prepare_conditions_cached(cond["xf_out"], xf_proj, self._condition(model_kwargs.get("c")))
        return eng

    def forward_test(self, h=None, src_mask=None, emb=None, xf_out=None, c=None, **kwargs):
        """controlnet_mcm.py:306-361: base / control blocks and `out` on an already embedded h; `c` is the control
        condition (raw or pre-encoded), None = the plain base model."""
        if h is None or emb is None or xf_out is None:
            raise McmError("forward_test needs h, emb and xf_out (controlnet_mcm.py:306-361)")
        eng = self.engine(h.shape[0])
        z = getattr(self, "_ft_zero", None)
        if z is None or z.shape[0] != h.shape[0] or z.device != h.device:
            self._ft_zero = z = torch.zeros(h.shape[0], self.base_model.time_embed_dim, device=h.device)
        eng.prepare_conditions_cached(xf_out, z, self._condition(c))
        return eng.layers_forward(h, emb)

    def forward(self, motion, timesteps, motion_mask=None, motion_length=None, num_intervals=1, c=None, **kwargs):
        """controlnet_mcm.py:168-233 + forward_test :306-361 (eval)."""
        kwargs = dict(kwargs)
        kwargs["c"] = c
        eng = self.bind_for_sampling(motion.shape[0], kwargs, motion.device)
        return eng.denoise(motion, timesteps)


SUBMODULES.register_module(module=ControlT2MHalf_MCM)


def ctrl_state_shapes(seq_len, n_ctrl, cond_feats, latent_dim=512):
    """name -> shape of a ControlT2MHalf_MCM state_dict (condition_pre_encode=False):
    base_model.* + controlnet.{j}.{copied_block.*, before_proj.*, after_proj.*} + control_cond_input.*
    (controlnet_mcm.py:107-153)."""
    base = state_shapes(seq_len=seq_len, latent_dim=latent_dim)
    out = {"base_model." + k: v for k, v in base.items()}
    p0 = "temporal_decoder_blocks.0."
    for j in range(n_ctrl):
        for k, v in base.items():
            if k.startswith(p0):
                out[f"controlnet.{j}.copied_block." + k[len(p0):]] = v
        if j == 0:
            out["controlnet.0.before_proj.weight"] = (latent_dim, latent_dim)
            out["controlnet.0.before_proj.bias"] = (latent_dim,)
        out[f"controlnet.{j}.after_proj.weight"] = (latent_dim, latent_dim)
        out[f"controlnet.{j}.after_proj.bias"] = (latent_dim,)
    out["control_cond_input.weight"] = (latent_dim, cond_feats)
    out["control_cond_input.bias"] = (latent_dim,)
    return out


def engine_state_from_ctrl(sd):
    """ControlT2MHalf_MCM state_dict -> the flat key space of the C-ABI (base_model. prefix stripped,
    dead ffn_channel weights dropped)."""
    out = {}
    for k, v in sd.items():
        if ".ffn_channel." in k:
            continue
        out[k[len("base_model."):] if k.startswith("base_model.") else k] = v
    return out
