"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pieces of Path B (configs/stmogen/*, SURVEY.md section 8 rows b1, b3, b7,
b8, b9) that the reference can still pin in this image (the `tutel` MoE of row b2 is an un-vendored, unpinned third-party
dependency: **parity unpinned**, not restated).  Each function cites the reference lines it follows; validated bit for bit
against the unmodified reference modules in oracle/make_golden.py::pathb (goldens in tests/golden/pathb.npz)."""
import torch
import torch.nn.functional as F

PART_ORDER = ("head", "stem", "larm", "rarm", "lleg", "rleg", "root", "trans", "face", "lhand", "rhand")   # stmogen.py:344-349


def smplx_slice(name):
    """get_smplx_slice, mogen/models/transformers/stmogen.py:53-68: columns of the 322-dim SMPL-X vector per body part."""
    j = lambda k: [k * 3, k * 3 + 1, k * 3 + 2]  # noqa: E731
    table = {
        "root": [0, 1, 2] + list(range(312, 322)),
        "trans": [309, 310, 311],
        "head": j(12) + j(15) + [156, 157, 158],
        "stem": j(3) + j(6) + j(9),
        "larm": j(14) + j(17) + j(19) + j(21),
        "rarm": j(13) + j(16) + j(18) + j(20),
        "lleg": j(2) + j(5) + j(8) + j(11),
        "rleg": j(1) + j(4) + j(7) + j(10),
        "face": list(range(159, 309)),
        "lhand": list(range(66, 111)),
        "rhand": list(range(111, 156)),
    }
    return table[name]


def body_slice():
    """PoseEncoder.body_slice for motionx (stmogen.py:306-308): the concatenation of the eleven part slices."""
    out = []
    for n in PART_ORDER:
        out.extend(smplx_slice(n))
    return out


def pose_encode(sd, motion, prefix="joint_embed."):
    """PoseEncoder.forward, motionx branch (stmogen.py:336-353, 376-378; gnn = identity): eleven gather + Linear(len(slice) ->
    L) and the whole-body Linear(322 -> L), concatenated to (B, T, 12 L)."""
    feats = [F.linear(motion[:, :, smplx_slice(n)].contiguous(), sd[f"{prefix}{n}_embed.weight"], sd[f"{prefix}{n}_embed.bias"])
             for n in PART_ORDER]
    feats.append(F.linear(motion[:, :, body_slice()].contiguous(), sd[prefix + "body_embed.weight"], sd[prefix + "body_embed.bias"]))
    return torch.cat(feats, dim=-1)


def pose_decode(sd, h, prefix="out.", out_dim=322):
    """PoseDecoder.forward, motionx branch with patch_size = 1 (stmogen.py:505-544): eleven Linear(L -> len(slice)) scattered
    into the 322 columns, the whole-body Linear(L -> 322) in body_slice order, output = (scatter + body) / 2."""
    B, T, _ = h.shape
    L = sd[prefix + "head_out.weight"].shape[1]
    output = torch.zeros(B, T, out_dim, dtype=h.dtype)
    for i, n in enumerate(PART_ORDER):
        output[:, :, smplx_slice(n)] = F.linear(h[:, :, i * L:(i + 1) * L].contiguous(), sd[f"{prefix}{n}_out.weight"],
                                                sd[f"{prefix}{n}_out.bias"])
    body = F.linear(h[:, :, 11 * L:].contiguous(), sd[prefix + "body_out.weight"], sd[prefix + "body_out.bias"])
    return (output + body) / 2.0


def static_body_mix(body_weight, body_value):
    """STMA.forward static branch (st_attention.py:123-128): softmax of the learned (H, H) human-topology graph over its
    second index, then every part is a mixture of all parts' value vectors.  body_value (B, T, H, L)."""
    w = F.softmax(body_weight, dim=1)
    return torch.einsum("hl,bnld->bnhd", w, body_value)


def scale_func(timestep, scale):
    """STMoGenTransformer.scale_func (stmogen.py:655-659), in Python floats like the reference."""
    w = (1 - (1000 - timestep) / 1000) * scale + 1
    return w, 1 - w


def cfg_combine(out_text, out_none, timestep, scale):
    """stmogen.py:755-759: out_text * text_coef + out_none * none_coef."""
    wt, wn = scale_func(int(timestep), scale)
    return out_text * wt + out_none * wn


def sffn(sd, x, emb, num_heads, prefix=""):
    """SFFN.forward (stmogen.py:596-607) with its StylizationBlock (stylization_block.py:29-40), inference (dropout = identity):
    per body part Linear -> GELU -> Linear, concatenated, then x + out_layers(LN(y) (1 + scale) + shift)."""
    B, T, D = x.shape
    xs = x.reshape(B, T, num_heads, -1)
    outs = []
    for i in range(num_heads):
        feat = xs[:, :, i].contiguous()
        feat = F.gelu(F.linear(feat, sd[f"{prefix}linear1_list.{i}.weight"], sd[f"{prefix}linear1_list.{i}.bias"]))
        outs.append(F.linear(feat, sd[f"{prefix}linear2_list.{i}.weight"], sd[f"{prefix}linear2_list.{i}.bias"]))
    y = torch.cat(outs, dim=-1)
    emb_out = F.linear(F.silu(emb), sd[prefix + "proj_out.emb_layers.1.weight"], sd[prefix + "proj_out.emb_layers.1.bias"]).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = F.layer_norm(y, (D,), sd[prefix + "proj_out.norm.weight"], sd[prefix + "proj_out.norm.bias"]) * (1 + scale) + shift
    h = F.linear(F.silu(h), sd[prefix + "proj_out.out_layers.2.weight"], sd[prefix + "proj_out.out_layers.2.bias"])
    return x.reshape(B, T, D) + h


def _stylization(sd, prefix, h, emb):
    """StylizationBlock.forward (stylization_block.py:29-40), inference."""
    D = h.shape[-1]
    emb_out = F.linear(F.silu(emb), sd[prefix + "emb_layers.1.weight"], sd[prefix + "emb_layers.1.bias"]).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = F.layer_norm(h, (D,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]) * (1 + scale) + shift
    return F.linear(F.silu(h), sd[prefix + "out_layers.2.weight"], sd[prefix + "out_layers.2.bias"])


def _body_self_attention(sd, prefix, x, num_heads=8):
    """EfficientSelfAttention.forward with time_embed_dim=None and an all-ones mask (efficient_attention.py:25-46), as STMA calls
    it on the (B*T, H, L) part tokens (st_attention.py:130-133)."""
    B, T, D = x.shape
    xn = F.layer_norm(x, (D,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    query = F.linear(xn, sd[prefix + "query.weight"], sd[prefix + "query.bias"])
    key = F.linear(xn, sd[prefix + "key.weight"], sd[prefix + "key.bias"]) + (1 - torch.ones(B, T, 1)) * -1000000
    query = F.softmax(query.view(B, T, num_heads, -1), dim=-1)
    key = F.softmax(key.view(B, T, num_heads, -1), dim=1)
    value = (F.linear(xn, sd[prefix + "value.weight"], sd[prefix + "value.bias"]) * torch.ones(B, T, 1)).view(B, T, num_heads, -1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    y = torch.einsum("bnhd,bhdl->bnhl", query, attention).reshape(B, T, D)
    return x + y


def stma_tail(sd, x, motion_feat, text_feat, emb, src_mask, cond_type, num_heads, latent_dim, static_body=True, dynamic_body=False,
              prefix=""):
    """STMA.forward (st_attention.py:105-175) from the point where both mixture-of-experts outputs exist: `motion_feat` =
    motion_moe(norm(x)) (B, T, H, 4L) and `text_feat` = text_moe(text_norm(xf)) (B, Nt, Ht, 2L) are INPUTS (the tutel MoE is
    un-vendored: parity unpinned, not restated)."""
    B, T, D = x.shape
    H, L = num_heads, latent_dim
    N = text_feat.shape[1] + T
    body_weight = F.softmax(sd[prefix + "body_weight"], dim=1)
    body_value = motion_feat[:, :, :, :L]
    body_feat = body_value
    if static_body:
        body_feat = torch.einsum("hl,bnld->bnhd", body_weight, body_value)
    body_feat = body_feat.reshape(B, T, D)
    if dynamic_body:
        body_feat = body_feat + _body_self_attention(sd, prefix + "body_d_attn.", body_value.reshape(B * T, H, -1)).reshape(B, T, D)
    text_cond_type = (cond_type % 10 > 0).float().unsqueeze(-1)
    src_mask = src_mask.view(B, T, 1, 1)
    key_text = text_feat[:, :, :, :L].contiguous()
    key_text = key_text + (1 - text_cond_type) * -1000000
    if text_feat.shape[2] == 1:
        key_text = key_text.repeat(1, 1, H, 1)
    key_motion = motion_feat[:, :, :, L:2 * L].contiguous()
    key_motion = key_motion + (1 - src_mask) * -1000000
    key = F.softmax(torch.cat((key_text, key_motion), dim=1).view(B, N, H, -1), dim=1)
    value_text = text_feat[:, :, :, L:].contiguous() * text_cond_type
    if text_feat.shape[2] == 1:
        value_text = value_text.repeat(1, 1, H, 1)
    value_motion = motion_feat[:, :, :, 2 * L:3 * L].contiguous() * src_mask
    value = torch.cat((value_text, value_motion), dim=1).view(B, N, H, -1)
    query = F.softmax(motion_feat[:, :, :, 3 * L:].contiguous().view(B, T, H, -1), dim=-1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    y_t = torch.einsum("bnhd,bhdl->bnhl", query, attention).reshape(B, T, D)
    return x.reshape(B, T, D) + _stylization(sd, prefix + "proj_out.", body_feat + y_t, emb)
