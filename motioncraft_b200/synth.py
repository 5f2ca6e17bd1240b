"""Deterministic synthetic weights / inputs for parity tests and benchmarks.

There is no network for checkpoints or datasets, so BASELINE.json's workloads run on random-init
weights of the reference architecture and synthetic inputs of the reference shapes (SURVEY.md
section 8d).  Everything is drawn on the CPU from `torch.Generator`s seeded per tensor NAME (crc32),
so the values do not depend on iteration order, device or world size, and the CPU oracle and the
GPU path see identical bits.

The reference zero-initialises `out`, every `StylizationBlock.out_layers[2]` and every
`FFN.linear2` (stylization_block.py:26, diffusion_transformer.py:20,97); a zero-init model outputs
exactly 0, so ALL parameters are re-drawn here with fan-in scaling to make every block live.
"""
import zlib

import torch

# seeds fixed by SURVEY.md section 8(d)
SEED_WEIGHTS, SEED_XT, SEED_XF_OUT, SEED_XF_PROJ, SEED_STEP_NOISE, SEED_C_EMB, SEED_C_M2D = 0, 123, 124, 125, 126, 127, 128
SEED_REPAINT_GT, SEED_REPAINT_NOISE = 131, 132
SEED_CLIP_FEAT = 133


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def synth_param(name, shape, seed=SEED_WEIGHTS):
    """One parameter tensor (fp32, CPU) as a pure function of (name, shape, seed)."""
    g = _gen(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = ".norm." in name or name.endswith("norm.weight") or name.endswith("norm.bias") or "text_norm" in name
    if name.endswith("sequence_embedding"):
        return torch.randn(shape, generator=g)
    if leaf == "running_var":                      # BatchNorm statistics of the WavEncoder: positive
        return 0.5 + torch.rand(shape, generator=g)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if ".bn" in name or ".downsample.1." in name:  # BatchNorm affine
        return (1.0 if leaf == "weight" else 0.0) + 0.1 * torch.randn(shape, generator=g)
    if is_norm:
        if leaf == "weight":
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "weight" and len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        # `out` reads a residual stream that has accumulated 24 O(1) block outputs; the 0.2 gain
        # keeps the predicted eps at O(1) rms so the 50-step trajectory stays in a sane range.
        gain = 0.2 if (name == "out.weight" or name.endswith(".out.weight")) else 1.0
        return torch.randn(shape, generator=g) * (gain * fan_in ** -0.5)
    return 0.02 * torch.randn(shape, generator=g)


def synth_state_dict(shapes, seed=SEED_WEIGHTS):
    """shapes: mapping name -> shape (e.g. {k: v.shape for k, v in module.state_dict().items()})."""
    return {k: synth_param(k, s, seed) for k, s in shapes.items()}


def synth_tensor(name, shape, seed):
    """Global (un-sharded) N(0,1) input tensor; multi-GPU ranks slice rows of this."""
    return torch.randn(tuple(shape), generator=_gen(seed, name))


def synth_rows(name, row_shape, seed, start, stop):
    """Rows [start, stop) of a virtual (N, *row_shape) N(0,1) tensor, each row seeded by its global
    index, so a rank can materialise only its slice and 1-GPU / N-GPU runs see identical rows."""
    rows = [torch.randn(tuple(row_shape), generator=_gen(seed, f"{name}[{i}]")) for i in range(start, stop)]
    return torch.stack(rows, 0) if rows else torch.empty((0,) + tuple(row_shape))
