"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (`/root/reference`) under a sys.modules shim.

The reference cannot be imported as-is in this image (mmcv, clip, smplx, ... are absent and
`mogen/__init__.py:46-54` asserts an mmcv version).  This module pre-registers empty package
objects for `mogen`, `mogen.models`, `mogen.models.{attentions,transformers,utils}` whose
`__path__` points into the reference tree (so the reference's own `__init__.py` files, which
star-import every model family, never run) and injects minimal stand-ins for
`mmcv.{cnn.MODELS, utils.Registry, runner.BaseModule}` and `clip`.  After `install()` the hot-path
files import unmodified:

    mogen.models.transformers.mcm                (MCMTransformer, DecoderLayer)
    mogen.models.transformers.controlnet_mcm     (ControlT2MHalf_MCM)
    mogen.models.attentions.efficient_attention  (EfficientSelfAttention, EfficientCrossAttention)
    mogen.models.utils.gaussian_diffusion        (GaussianDiffusion, SpacedDiffusion, space_timesteps)

Only `oracle/make_golden.py`, `oracle/validate_oracle.py` and `bench.py --impl reference` (when the
reference tree is present) use this.  It reads `/root/reference` and therefore only works in the
build container; nothing on the GPU box may depend on it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MCM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mogen", "models"))


class _Registry:
    """Just enough of mmcv.utils.Registry for `mogen/models/builder.py:1-36`."""

    def __init__(self, name, parent=None, build_func=None, **kw):
        self.name = name
        self.parent = parent
        self._module_dict = {}
        self.build_func = build_func or _default_build

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, cfg, *a, **kw):
        return self.build_func(cfg, self, *a, **kw)


def _default_build(cfg, registry, default_args=None):
    if cfg is None:
        return None
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


SOT_TOKEN, EOT_TOKEN = 49406, 49407     # CLIP's start / end-of-text ids; EOT is the largest id, which is what
                                        # `text.argmax(dim=-1)` (diffusion_transformer.py:165) relies on


def stub_clip_tokenize(texts, context_length=77, truncate=False):
    """Stand-in for clip.tokenize (the BPE vocabulary is not available offline): one pseudo-token per whitespace-separated
    word, framed by SOT / EOT and zero-padded to 77 -- enough to reproduce the ONE thing the trainable text stack takes
    from the tokens, the position of EOT."""
    import torch
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int)
    for i, t in enumerate(texts):
        words = t.split()[: context_length - 2]
        ids = [SOT_TOKEN] + [1000 + (sum(map(ord, w)) % 40000) for w in words] + [EOT_TOKEN]
        out[i, : len(ids)] = torch.tensor(ids, dtype=torch.int)
    return out


def stub_clip_load(name, device="cpu", **kw):
    """Stand-in for clip.load: an object with the attributes `encode_text(..., clip_feat=...)` touches (`.dtype`);
    the frozen CLIP tower itself is not available, so `clip_feat` must be supplied."""
    import torch
    import torch.nn as nn

    class _FrozenClip(nn.Module):
        dtype = torch.float32

    return _FrozenClip(), None


def install():
    """Idempotently install the shim; returns the reference root."""
    if "mogen" in sys.modules and getattr(sys.modules["mogen"], "_mcm_shim", False):
        return REFERENCE_ROOT
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch.nn as nn

    mmcv = types.ModuleType("mmcv")
    mmcv.__version__ = "1.7.0"
    mmcv_cnn = types.ModuleType("mmcv.cnn")
    mmcv_utils = types.ModuleType("mmcv.utils")
    mmcv_runner = types.ModuleType("mmcv.runner")
    root = _Registry("model")
    mmcv_cnn.MODELS = root
    mmcv_utils.Registry = _Registry
    mmcv_utils.build_from_cfg = _default_build

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

        def init_weights(self):
            pass

    mmcv_runner.BaseModule = BaseModule
    mmcv.cnn, mmcv.utils, mmcv.runner = mmcv_cnn, mmcv_utils, mmcv_runner
    for name, mod in (("mmcv", mmcv), ("mmcv.cnn", mmcv_cnn), ("mmcv.utils", mmcv_utils),
                      ("mmcv.runner", mmcv_runner)):
        sys.modules.setdefault(name, mod)

    clip = types.ModuleType("clip")
    clip.load = stub_clip_load
    clip.tokenize = stub_clip_tokenize
    sys.modules.setdefault("clip", clip)

    def _pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, *rel.split("/"))]
        m._mcm_shim = True
        sys.modules[name] = m
        return m

    _pkg("mogen", "mogen")
    _pkg("mogen.models", "mogen/models")
    _pkg("mogen.models.attentions", "mogen/models/attentions")
    _pkg("mogen.models.transformers", "mogen/models/transformers")
    _pkg("mogen.models.utils", "mogen/models/utils")
    _pkg("mogen.models.gnns", "mogen/models/gnns")
    # mogen/models/gnns/stgcn.py (imported by stmogen.py) asks mmcv.cnn for layer factories at import / build time
    mmcv_cnn.build_norm_layer = lambda cfg, n, postfix="": ("bn" + str(postfix), nn.BatchNorm2d(n))
    mmcv_cnn.build_activation_layer = lambda cfg: nn.ReLU()
    return REFERENCE_ROOT


TEXT_ENCODER_CFG = dict(pretrained_model="clip", latent_dim=256, num_layers=4, num_heads=4, ff_size=2048, dropout=0,
                        use_text_proj=True)      # configs/mcm/mcm_t2m_smplx.py:58-64


def build_reference_mcm(T=196, num_layers=8, input_feats=322, latent_dim=512, time_embed_dim=2048,
                        text_latent_dim=256, ff_size=1024, num_heads=4, text_encoder=None):
    """Reference MCMTransformer built exactly as configs/mcm/mcm_t2m_smplx.py:37-58 does, minus CLIP
    (text_encoder=TEXT_ENCODER_CFG builds the trainable text stack around the clip stand-in)."""
    install()
    import mogen.models.attentions.efficient_attention  # noqa: F401  (registers the attention types)
    from mogen.models.transformers.mcm import MCMTransformer
    m = MCMTransformer(
        input_feats=input_feats, max_seq_len=T, latent_dim=latent_dim, time_embed_dim=time_embed_dim,
        num_layers=num_layers,
        sa_block_cfg=dict(type="EfficientSelfAttention", latent_dim=T, num_heads=num_heads, dropout=0,
                          time_embed_dim=time_embed_dim),
        ca_block_cfg=dict(type="EfficientCrossAttention", latent_dim=latent_dim,
                          text_latent_dim=text_latent_dim, num_heads=num_heads, dropout=0,
                          time_embed_dim=time_embed_dim),
        ffn_cfg=dict(latent_dim=latent_dim, ffn_dim=ff_size, dropout=0, time_embed_dim=time_embed_dim),
        text_encoder=text_encoder)
    m.use_text_proj = True  # what text_encoder=dict(use_text_proj=True) sets (diffusion_transformer.py:117)
    return m.eval()


def reference_opt():
    """The argparse namespace tools/*.py smuggle into SpacedDiffusion."""
    from argparse import Namespace
    return Namespace(no_repaint=False, same_overlap_noisy=False, addBlend=True, overlap_len=0,
                     no_resample=False, timestep_respacing="ddim50", jump_length=3, jump_n_sample=5)


def build_reference_diffusion(respace=None, steps=1000):
    install()
    from mogen.models.utils import gaussian_diffusion as gd
    betas = gd.get_named_beta_schedule("linear", steps)
    kw = dict(betas=betas, model_mean_type=gd.ModelMeanType.EPSILON,
              model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    if respace is not None:
        return gd.SpacedDiffusion(use_timesteps=gd.space_timesteps(steps, respace), opt=reference_opt(), **kw)
    return gd.GaussianDiffusion(**kw)
