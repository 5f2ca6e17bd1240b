"""First pieces of Path B -- the STMoGen family of configs/stmogen/* (SURVEY.md section 8 rows b1, b3, b8, b9 and the start_x /
fixed_large sampler parameterisation), built on the same CUDA library as the configs/mcm path.

What is here runs on the device and is pinned against the unmodified reference (tests/golden/pathb.npz):

  PoseEncoder / PoseDecoder  stmogen.py:141-578 (motionx, 12 parts): the reference gathers eleven column slices of the 322-dim
      vector, runs eleven small Linears plus a whole-body Linear and concatenates / scatters.  Here the twelve weight matrices
      are laid out once as ONE block-structured matrix -- (12 L x 322) for the encoder, (322 x 12 L) with the
      `(scatter + body) / 2` folded in for the decoder; the slices partition the 322 columns, so the structure is exact --
      and the whole layer is one pass of the tcgen05 GEMM kernel in its 3-pass bf16-split mode (the precision class of
      joint_embed / out).  Gather and scatter cost nothing: they are zero blocks of the operand.
  SFFN                       stmogen.py:581-607 (+ its StylizationBlock): the twelve per-part Linear -> GELU -> Linear pairs as TWO
      block-diagonal tcgen05 GEMM launches (batch = body part, per-part biases), LayerNorm + AdaLN + SiLU, output Linear with the
      residual in its epilogue -> `mcm_sffn_forward`
  STMATail                   st_attention.py:105-175 minus the two MoE layers: static + dynamic body branches, the temporal linear
      attention over text + motion tokens with the reference's masks, StylizationBlock -> `mcm_stma_mix` (MoE outputs are inputs)
  static_body_mix            st_attention.py:123-128 -> `mcm_part_mix`
  cfg_combine / scale_func   stmogen.py:655-659, 755-759 -> `mcm_cfg_combine`
  start_x / fixed_large      diffusion.py (`SamplerTables(model_mean="start_x")`)

NOT here: the mixture-of-experts of STMA (`tutel.moe.moe_layer`, st_attention.py:17-56) -- an un-vendored, unpinned
third-party dependency whose routing cannot be pinned in this image -- and therefore STMoGenTransformer itself; the
dynamic / temporal branches follow once the MoE question is settled.  `STMoGenTransformer` raises accordingly.
"""
import ctypes

import torch
from torch import nn

from . import _lib
from ._lib import McmError
from .engine import test_linear

PART_ORDER = ("head", "stem", "larm", "rarm", "lleg", "rleg", "root", "trans", "face", "lhand", "rhand")


def get_smplx_slice(name):
    """stmogen.py:53-68."""
    j = lambda k: [k * 3, k * 3 + 1, k * 3 + 2]  # noqa: E731
    return {
        "root": [0, 1, 2] + list(range(312, 322)), "trans": [309, 310, 311],
        "head": j(12) + j(15) + [156, 157, 158], "stem": j(3) + j(6) + j(9),
        "larm": j(14) + j(17) + j(19) + j(21), "rarm": j(13) + j(16) + j(18) + j(20),
        "lleg": j(2) + j(5) + j(8) + j(11), "rleg": j(1) + j(4) + j(7) + j(10),
        "face": list(range(159, 309)), "lhand": list(range(66, 111)), "rhand": list(range(111, 156)),
    }[str(name)]


def _body_slice():
    out = []
    for n in PART_ORDER:
        out.extend(get_smplx_slice(n))
    return out


class _PackedLinear(nn.Module):
    """Parameter container whose forward is one tcgen05 GEMM over a block-structured weight assembled from its Linears."""

    def __init__(self):
        super().__init__()
        self._packed = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: setattr(self, "_packed", None))

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _gemm(self, x, n_out):
        if x.device.type != "cuda":
            raise McmError("motioncraft_b200 modules compute on an sm_100a CUDA device only (no CPU fallback)")
        if self._packed is None or self._packed[0].device != x.device:
            with torch.no_grad():
                self._packed = self._pack(x.device)
        W, b = self._packed
        lead = x.shape[:-1]
        y = test_linear(x.reshape(-1, x.shape[-1]), W, b, fmt=1)          # fmt 1 = bf16 hi/lo split, 3 passes
        return y.view(*lead, n_out)


class PoseEncoder(_PackedLinear):
    """stmogen.py:141-378, `dataset_name='motionx'`, joints=False, body_graph=False (what configs/stmogen/* build)."""

    def __init__(self, dataset_name="motionx", latent_dim=128, input_dim=322, patch_size=1, joints=False, body_graph=False,
                 gnn_cfg=None):
        super().__init__()
        if dataset_name != "motionx" or joints or body_graph or patch_size != 1 or input_dim != 322:
            raise McmError("PoseEncoder: the 12-part motionx layout (patch_size 1, no joint / graph variants) is implemented")
        self.dataset_name, self.latent_dim, self.parts_num = dataset_name, latent_dim, 12
        for n in PART_ORDER:
            setattr(self, n + "_embed", nn.Linear(len(get_smplx_slice(n)), latent_dim))
        self.body_embed = nn.Linear(input_dim, latent_dim)

    def _pack(self, dev):
        L = self.latent_dim
        W = torch.zeros(12 * L, 322, device=dev)
        b = torch.empty(12 * L, device=dev)
        for i, n in enumerate(PART_ORDER):
            lin = getattr(self, n + "_embed")
            W[i * L:(i + 1) * L, get_smplx_slice(n)] = lin.weight.to(dev)
            b[i * L:(i + 1) * L] = lin.bias.to(dev)
        W[11 * L:, _body_slice()] = self.body_embed.weight.to(dev)
        b[11 * L:] = self.body_embed.bias.to(dev)
        return W, b

    def forward(self, motion):
        return self._gemm(motion, 12 * self.latent_dim)


class PoseDecoder(_PackedLinear):
    """stmogen.py:380-578, motionx, patch_size 1: output = (scatter(part outputs) + body_out(h_body)) / 2."""

    def __init__(self, dataset_name="motionx", latent_dim=128, output_dim=322, patch_size=1, joints=False):
        super().__init__()
        if dataset_name != "motionx" or joints or patch_size != 1 or output_dim != 322:
            raise McmError("PoseDecoder: the 12-part motionx layout (patch_size 1, no joint variant) is implemented")
        self.dataset_name, self.latent_dim, self.output_dim = dataset_name, latent_dim, output_dim
        for n in PART_ORDER:
            setattr(self, n + "_out", nn.Linear(latent_dim, len(get_smplx_slice(n))))
        self.body_out = nn.Linear(latent_dim, output_dim)

    def _pack(self, dev):
        L = self.latent_dim
        W = torch.zeros(322, 12 * L, device=dev)
        b = torch.zeros(322, device=dev)
        for i, n in enumerate(PART_ORDER):
            lin = getattr(self, n + "_out")
            cols = get_smplx_slice(n)
            W[cols, i * L:(i + 1) * L] = lin.weight.to(dev)
            b[cols] = lin.bias.to(dev)
        # body_out's output k is added to column k un-permuted (stmogen.py:516, 543); only the ENCODER's whole-body Linear
        # reads its input through body_slice
        W[:, 11 * L:] = self.body_out.weight.to(dev)
        b = b + self.body_out.bias.to(dev)
        return 0.5 * W, 0.5 * b            # (scatter + body) / 2: a power of two, exact in every operand format

    def forward(self, h):
        return self._gemm(h, self.output_dim)


def static_body_mix(body_weight, body_value):
    """STMA static branch (st_attention.py:123-128): body_value (B, T, H, L) fp32 CUDA -> same shape."""
    if body_value.device.type != "cuda":
        raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only")
    v = body_value.detach().float().contiguous()
    w = body_weight.detach().to(v.device, torch.float32).contiguous()
    H, L = v.shape[-2], v.shape[-1]
    out = torch.empty_like(v)
    lib = _lib.load()
    with torch.cuda.device(v.device):
        _lib.check(lib.mcm_part_mix(ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                    v.numel() // (H * L), H, L, ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)))
    return out


def scale_func(timestep, scale=6.5):
    """STMoGenTransformer.scale_func (stmogen.py:655-659)."""
    w = (1 - (1000 - timestep) / 1000) * scale + 1
    return {"text_coef": w, "none_coef": 1 - w}


def cfg_combine(out_text, out_none, timestep, scale=6.5):
    """stmogen.py:755-759 on the device; `timestep` is the Python int the caller already holds (the reference reads
    `int(timesteps[0])` from the device every step -- a host synchronisation this signature avoids)."""
    if out_text.device.type != "cuda":
        raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only")
    a, b = out_text.detach().float().contiguous(), out_none.detach().float().contiguous()
    coef = scale_func(int(timestep), scale)
    out = torch.empty_like(a)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.mcm_cfg_combine(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), float(coef["text_coef"]),
                                       float(coef["none_coef"]), ctypes.c_void_p(out.data_ptr()), a.numel(),
                                       ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))
    return out


class _Stylization(nn.Module):
    """Parameter container with the reference's StylizationBlock key names (stylization_block.py:14-27)."""

    def __init__(self, latent_dim, time_embed_dim):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * latent_dim))
        self.norm = nn.LayerNorm(latent_dim)
        self.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=0.0), nn.Linear(latent_dim, latent_dim))


class SFFN(nn.Module):
    """stmogen.py:581-607 -- same constructor arguments and state_dict keys; forward runs on the device in `mcm_sffn_forward`
    (inference: dropout is the identity).  x (B, T, num_heads * latent_dim), emb (B, time_embed_dim)."""

    def __init__(self, latent_dim, ffn_dim, dropout, time_embed_dim, **kwargs):
        super().__init__()
        self.num_heads = kwargs["num_heads"]
        self.latent_dim, self.ffn_dim, self.time_embed_dim = latent_dim, ffn_dim, time_embed_dim
        self.linear1_list = nn.ModuleList(nn.Linear(latent_dim, ffn_dim) for _ in range(self.num_heads))
        self.linear2_list = nn.ModuleList(nn.Linear(ffn_dim, latent_dim) for _ in range(self.num_heads))
        self.proj_out = _Stylization(latent_dim * self.num_heads, time_embed_dim)
        self._stacked = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: setattr(self, "_stacked", None))

    def _apply(self, fn, *a, **k):
        self._stacked = None
        return super()._apply(fn, *a, **k)

    def _stack(self, dev):
        if self._stacked is None or self._stacked[0].device != dev:
            f = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
            po = self.proj_out
            self._stacked = tuple(f(t) for t in (
                torch.stack([m.weight for m in self.linear1_list]), torch.stack([m.bias for m in self.linear1_list]),
                torch.stack([m.weight for m in self.linear2_list]), torch.stack([m.bias for m in self.linear2_list]),
                po.emb_layers[1].weight, po.emb_layers[1].bias, po.norm.weight, po.norm.bias,
                po.out_layers[2].weight, po.out_layers[2].bias))
        return self._stacked

    def forward(self, x, emb, **kwargs):
        if x.device.type != "cuda":
            raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only")
        B, T, D = x.shape
        if D != self.num_heads * self.latent_dim:
            raise McmError(f"SFFN: x has {D} features, expected {self.num_heads * self.latent_dim}")
        xc = x.detach().float().contiguous()
        ec = emb.detach().to(x.device, torch.float32).contiguous()
        out = torch.empty_like(xc)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.mcm_sffn_forward(B, T, self.num_heads, self.latent_dim, self.ffn_dim, self.time_embed_dim, ptr(xc), ptr(ec),
                                            *[ptr(t) for t in self._stack(x.device)], ptr(out),
                                            ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out


class _BodyAttention(nn.Module):
    """Parameter container with the key names of EfficientSelfAttention(latent_dim, 8 heads, time_embed_dim=None)
    (efficient_attention.py:12-23), the dynamic body branch of STMA (st_attention.py:88-93)."""

    def __init__(self, latent_dim):
        super().__init__()
        self.norm = nn.LayerNorm(latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(latent_dim, latent_dim)
        self.value = nn.Linear(latent_dim, latent_dim)


class STMATail(nn.Module):
    """Everything of STMA (st_attention.py:65-175) except its two mixture-of-experts layers: the parameters `body_weight`,
    `body_d_attn.*`, `proj_out.*` under the reference's key names, and `forward` = `mcm_stma_mix` on the MoE OUTPUTS
    (motion_feat (B, T, H, 4L), text_feat (B, Nt, Ht, 2L)).  With a pinned MoE in front of it this is STMA.forward."""

    def __init__(self, latent_dim, num_heads, num_text_heads, time_embed_dim, static_body=True, dynamic_body=False, **kwargs):
        super().__init__()
        self.latent_dim, self.num_heads, self.num_text_heads = latent_dim, num_heads, num_text_heads
        self.time_embed_dim, self.static_body, self.dynamic_body = time_embed_dim, static_body, dynamic_body
        self.body_weight = nn.Parameter(torch.randn(num_heads, num_heads))
        if dynamic_body:
            self.body_d_attn = _BodyAttention(latent_dim)
        self.proj_out = _Stylization(latent_dim * num_heads, time_embed_dim)
        self._stacked = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: setattr(self, "_stacked", None))

    def _apply(self, fn, *a, **k):
        self._stacked = None
        return super()._apply(fn, *a, **k)

    def _stack(self, dev):
        if self._stacked is None or self._stacked[0].device != dev:
            f = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
            po = self.proj_out
            dyn = [None] * 4
            if self.dynamic_body:
                a = self.body_d_attn
                dyn = [f(a.norm.weight), f(a.norm.bias), f(torch.cat([a.query.weight, a.key.weight, a.value.weight])),
                       f(torch.cat([a.query.bias, a.key.bias, a.value.bias]))]
            self._stacked = (f(self.body_weight), *dyn, f(po.emb_layers[1].weight), f(po.emb_layers[1].bias), f(po.norm.weight),
                             f(po.norm.bias), f(po.out_layers[2].weight), f(po.out_layers[2].bias))
        return self._stacked

    def forward(self, x, motion_feat, text_feat, emb, src_mask, cond_type, **kwargs):
        if x.device.type != "cuda":
            raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only")
        B, T, D = x.shape
        H, L = self.num_heads, self.latent_dim
        Nt, Ht = text_feat.shape[1], text_feat.shape[2]
        if D != H * L or tuple(motion_feat.shape) != (B, T, H, 4 * L) or tuple(text_feat.shape) != (B, Nt, Ht, 2 * L):
            raise McmError("STMATail: x (B, T, H*L), motion_feat (B, T, H, 4L), text_feat (B, Nt, Ht, 2L) expected")
        f = lambda t: t.detach().to(x.device, torch.float32).contiguous()  # noqa: E731
        xc, mf, tf, ec = f(x), f(motion_feat), f(text_feat), f(emb)
        mask = f(src_mask).reshape(B, T)
        tcond = (cond_type.reshape(B).to(x.device) % 10 > 0).float().contiguous()              # st_attention.py:137
        out = torch.empty_like(xc)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else None)  # noqa: E731
        st = self._stack(x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.mcm_stma_mix(B, T, H, L, Nt, Ht, self.time_embed_dim, 1 if self.static_body else 0, ptr(xc), ptr(mf), ptr(tf),
                                        ptr(ec), ptr(mask), ptr(tcond), *[ptr(t) for t in st], ptr(out),
                                        ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out


class STMoGenTransformer(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise McmError("STMoGenTransformer (configs/stmogen/*) is not complete in motioncraft_b200: its STMA blocks route "
                       "through tutel's mixture-of-experts layer, an un-vendored dependency whose semantics cannot be pinned "
                       "here (SURVEY.md section 8 row b2).  Available pieces: PoseEncoder, PoseDecoder, SFFN, STMATail (STMA after its MoE layers), static_body_mix, "
                       "cfg_combine, the start_x / fixed_large samplers (motioncraft_b200/pathb.py)")
