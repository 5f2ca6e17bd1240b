"""DenoiserEngine -- Python owner of one `mcm_ctx` (include/mcm_b200.h).

PyTorch is plumbing here: it owns device memory (parameters, inputs, outputs) and the stream; all
arithmetic of the hot path happens inside libmcm_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import McmConfig, McmError, McmSampler


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class SamplerTables:
    """float32 host copies of the (respaced) diffusion tables in the layout of `mcm_sampler`."""

    FIELDS = ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
              "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped")

    def __init__(self, tables, timestep_map, mode, eta=0.0, seed=0, model_mean="epsilon"):
        self.n_steps = len(timestep_map)
        self._tmap = np.ascontiguousarray(np.asarray(timestep_map, dtype=np.int32))
        # the float64 -> float32 cast _extract_into_tensor applies (gaussian_diffusion.py:1340)
        self._arrs = {k: np.ascontiguousarray(np.asarray(tables[k], dtype=np.float64).astype(np.float32))
                      for k in self.FIELDS}
        s = McmSampler()
        s.mode = {"ddim": 0, "ddpm": 1}[mode]
        s.n_steps = self.n_steps
        s.eta = float(eta)
        s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF      # key of the on-device noise of stochastic samplers (no explicit noise given)
        s.model_mean_type = {"epsilon": 0, "start_x": 1}[model_mean]
        s.timestep_map = self._tmap.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
        for k in self.FIELDS:
            setattr(s, k, self._arrs[k].ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        self.struct = s


class DenoiserEngine:
    def __init__(self, state_dict, *, seq_len, input_feats=322, latent_dim=512, time_embed_dim=2048, ffn_dim=1024,
                 text_latent_dim=256, num_heads=4, num_layers=8, num_ctrl_blocks=0, ctrl_cond_feats=0,
                 max_batch=1, max_text_tokens=77, precise_all=False, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise McmError("motioncraft_b200 runs on an sm_100a CUDA device only; there is no CPU path")
        cfg = McmConfig(input_feats, seq_len, latent_dim, time_embed_dim, ffn_dim, text_latent_dim, num_heads,
                        num_layers, num_ctrl_blocks, ctrl_cond_feats, max_batch, max_text_tokens, int(precise_all))
        self.cfg = cfg
        self.seq_len, self.input_feats, self.latent_dim = seq_len, input_feats, latent_dim
        self.time_embed_dim, self.max_batch = time_embed_dim, max_batch
        self._ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_create(ctypes.byref(cfg), ctypes.byref(self._ctx)))
        self._cond_key = None
        self._cond_refs = None
        self.load_params(state_dict)

    def load_params(self, state_dict):
        """(Re)load weights: hands every parameter to the library and packs them (mcm_finalize_params).  On an engine
        that already holds weights the packed copies are rewritten IN PLACE -- workspace, streams and captured graphs
        survive a load_state_dict; only the step-invariant condition work must be prepared again."""
        with torch.cuda.device(self.device):
            keep = []
            for name, t in state_dict.items():
                if not torch.is_floating_point(t):
                    continue
                t = _f32c(t, self.device)
                keep.append(t)
                _lib.check(self.lib.mcm_set_param(self._ctx, name.encode(), _ptr(t), t.numel()))
            _lib.check(self.lib.mcm_finalize_params(self._ctx, _stream(self.device)))
            torch.cuda.synchronize(self.device)
        del keep
        self._cond_key = None
        self._cond_refs = None

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.mcm_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """Scheduling knobs ("dual", "graph", "chunk"); results are bit-identical in every mode."""
        _lib.check(self.lib.mcm_set_option(self._ctx, name.encode(), int(value)))

    def debug_copy(self, what, numel):
        """Development aid: fp16 dump of the fused kernel's operand tile (what=0) or hidden scratch (what=1)."""
        out = torch.empty(numel, device=self.device, dtype=torch.float16)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_debug_copy(self._ctx, int(what), _ptr(out), out.numel() * 2))
        return out

    # ------------------------------------------------------------------ conditions
    def prepare_conditions(self, xf_out, xf_proj, c=None):
        xf_out = _f32c(xf_out, self.device)
        xf_proj = _f32c(xf_proj, self.device)
        B, N = xf_out.shape[0], xf_out.shape[1]
        c_len = 0
        if c is not None:
            c = _f32c(c, self.device)
            c_len = c.shape[1]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_prepare_conditions(self._ctx, B, _ptr(xf_out), N, _ptr(xf_proj), _ptr(c), c_len,
                                                       _stream(self.device)))
        self._keep_cond = (xf_out, xf_proj, c)

    def prepare_conditions_cached(self, xf_out, xf_proj, c=None):
        """prepare_conditions unless the SAME tensor objects (identity, version counter, shape) were prepared last.
        The cache entry holds strong references to the caller's tensors: their storage cannot be freed and handed to a
        new batch at the same address (a fresh tensor has _version 0) while the entry is alive."""
        key = tuple((id(t), t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None
                    for t in (xf_out, xf_proj, c))
        if key != self._cond_key:
            self.prepare_conditions(xf_out, xf_proj, c)
            self._cond_key = key
            self._cond_refs = (xf_out, xf_proj, c)

    def invalidate_conditions(self):
        self._cond_key = None
        self._cond_refs = None

    # ------------------------------------------------------------------ per-step / per-block
    def denoise(self, x, timesteps):
        x = _f32c(x, self.device)
        B = x.shape[0]
        out = torch.empty_like(x)
        t_uniform, t_dev = 0, None
        if isinstance(timesteps, int):
            t_uniform = int(timesteps)
        else:
            t_dev = timesteps.detach().to(device=self.device, dtype=torch.int64).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_denoise(self._ctx, B, _ptr(x), _ptr(t_dev), t_uniform, _ptr(out),
                                            _stream(self.device)))
        return out

    def block_forward(self, kind, index, x, emb):
        x = _f32c(x, self.device).clone()
        emb = _f32c(emb, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_block_forward(self._ctx, kind, index, x.shape[0], _ptr(x), _ptr(emb),
                                                  _stream(self.device)))
        return x

    def layers_forward(self, h, emb):
        """MCMTransformer.forward_test (mcm.py:93-102): decoder layers + `out` on an already embedded h."""
        h = _f32c(h, self.device)
        emb = _f32c(emb, self.device)
        out = torch.empty(h.shape[0], h.shape[1], self.input_feats, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_layers_forward(self._ctx, h.shape[0], _ptr(h), _ptr(emb), _ptr(out),
                                                   _stream(self.device)))
        return out

    # ------------------------------------------------------------------ sampler
    def sample(self, tables: SamplerTables, x_T, step_noise=None):
        """x_T on the device -> x_0 on the device (new tensor).  step_noise=None with a stochastic sampler: the
        library generates each step's noise on the device from tables.struct.seed."""
        x_T = _f32c(x_T, self.device)
        out = torch.empty_like(x_T)
        if step_noise is not None:
            step_noise = _f32c(step_noise, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_sample(self._ctx, ctypes.byref(tables.struct), x_T.shape[0], _ptr(x_T),
                                           _ptr(step_noise), _ptr(out), _stream(self.device)))
        return out

    def sample_repaint(self, tables: SamplerTables, x_T, gt, keep_mask, noise_seq, *, times=None, betas=None,
                       overlap_len=0, add_blend=True):
        """RePaint / outpainting DDIM sampling (mcm_sample_repaint).  gt / keep_mask broadcast to x_T's shape; `times` is
        the harmonising schedule (list ending with -1) or None for the plain loop; `noise_seq` [n_draws, B, T, F]."""
        x_T = _f32c(x_T, self.device)
        out = torch.empty_like(x_T)
        gt = _f32c(gt, self.device).expand_as(x_T).contiguous()
        keep = keep_mask.to(device=self.device, dtype=torch.bool).expand_as(x_T).contiguous().to(torch.uint8)
        if noise_seq is not None:
            noise_seq = _f32c(noise_seq, self.device)
        r = _lib.McmRepaint()
        keepalive = []
        if times is not None:
            t_arr = np.ascontiguousarray(np.asarray(times, dtype=np.int32))
            b_arr = np.ascontiguousarray(np.asarray(betas, dtype=np.float64).astype(np.float32))
            keepalive += [t_arr, b_arr]
            r.n_times = len(t_arr)
            r.times = t_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
            r.betas = b_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        else:
            r.n_times = 0
        r.gt, r.keep_mask = gt.data_ptr(), keep.data_ptr()
        r.noise_seq = noise_seq.data_ptr() if noise_seq is not None else None    # None: generated on the device
        r.n_draws = noise_seq.shape[0] if noise_seq is not None else 0
        r.overlap_len, r.add_blend = int(overlap_len), int(bool(add_blend))
        blend_w = torch.linspace(0, 1, int(overlap_len)).to(self.device) if overlap_len > 0 else None   # gaussian_diffusion.py:873
        r.blend_w = blend_w.data_ptr() if blend_w is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_sample_repaint(self._ctx, ctypes.byref(tables.struct), ctypes.byref(r), x_T.shape[0],
                                                   _ptr(x_T), _ptr(out), _stream(self.device)))
        # No host synchronisation: the host arrays (times, betas, tables) are read while the call ENQUEUES the loop, and
        # the device temporaries are released to torch's stream-ordered allocator on the stream the loop runs on.
        del keepalive
        return out

    def sample_host(self, tables: SamplerTables, x_T_host, out_host=None, step_noise_host=None):
        """x_T in (pinned) host memory -> x_0 in host memory; H2D + loop + D2H inside the library.  Optional
        step_noise_host [n_steps, B, T, F] in host memory is copied one step at a time."""
        assert x_T_host.device.type == "cpu" and x_T_host.dtype == torch.float32 and x_T_host.is_contiguous()
        if step_noise_host is not None:
            assert (step_noise_host.device.type == "cpu" and step_noise_host.dtype == torch.float32
                    and step_noise_host.is_contiguous() and step_noise_host.shape[0] == tables.n_steps)
        if out_host is None:
            out_host = torch.empty_like(x_T_host, pin_memory=x_T_host.is_pinned())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_sample_host(self._ctx, ctypes.byref(tables.struct), x_T_host.shape[0],
                                                _ptr(x_T_host), _ptr(step_noise_host), _ptr(out_host),
                                                _stream(self.device)))
        return out_host

    def test_linear(self, A, W, bias, fmt):
        A, W = _f32c(A, self.device), _f32c(W, self.device)
        bias = _f32c(bias, self.device) if bias is not None else None
        C = torch.empty(A.shape[0], W.shape[0], device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mcm_test_linear(A.shape[0], W.shape[0], A.shape[1], _ptr(A), _ptr(W), _ptr(bias),
                                                _ptr(C), fmt, _stream(self.device)))
        return C


def test_linear(A, W, bias=None, fmt=0):
    """Raw tcgen05 GEMM (kernel unit tests): C = A W^T + bias with fp16 (fmt 0) / bf16x2 (fmt 1) operands."""
    lib = _lib.load()
    dev = A.device
    A, W = _f32c(A, dev), _f32c(W, dev)
    bias = _f32c(bias, dev) if bias is not None else None
    C = torch.empty(A.shape[0], W.shape[0], device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.mcm_test_linear(A.shape[0], W.shape[0], A.shape[1], _ptr(A), _ptr(W), _ptr(bias), _ptr(C),
                                       int(fmt), _stream(dev)))
    return C
