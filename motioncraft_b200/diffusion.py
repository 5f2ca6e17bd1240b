"""Diffusion schedule + sampler front-end with the reference's class names and call signatures
(mogen/models/utils/gaussian_diffusion.py), re-hosted on the CUDA library.

The float64 numpy tables are computed exactly as `GaussianDiffusion.__init__` (:354-387) and
`SpacedDiffusion.__init__` (:1416-1431) do; the sampling LOOPS (`p_sample_loop`, `ddim_sample_loop`)
do not call a Python model per step: they hand the whole loop to `mcm_sample` (one C call, no host
synchronisation between steps).  The RePaint / outpainting branch of `ddim_sample_loop` (y['outpainting_mask'],
:855-884, :1050-1118) goes to `mcm_sample_repaint`.  Training-side methods (training_losses, VB terms) and
`opt.same_overlap_noisy` are out of scope and raise.
"""
import enum

import numpy as np
import torch

from ._lib import McmError
from .engine import SamplerTables
from .scheduler import count_draws, get_schedule_jump_cjm_ddim


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:235-253 (linear); the cosine schedule follows :256-271."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        import math
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        n = num_diffusion_timesteps
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """gaussian_diffusion.py:1346-1404.  Returns a set, as the reference does."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        if section_counts == "fast27":
            steps = space_timesteps(num_timesteps, "15,15,8,6,6")
            steps.remove(num_timesteps - 1)
            steps.add(num_timesteps - 3)
            return steps
        section_counts = [int(v) for v in section_counts.split(",")]
    n_sec = len(section_counts)
    base, extra = divmod(num_timesteps, n_sec)
    picked, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            picked.append(start + round(pos))
            pos += stride
        start += size
    return set(picked)


class GaussianDiffusion:
    """Schedule tables + on-device sampling loops.  Signature of gaussian_diffusion.py:319-344."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False, opt=None):
        if model_mean_type not in (ModelMeanType.EPSILON, ModelMeanType.START_X) or \
                model_var_type not in (ModelVarType.FIXED_SMALL, ModelVarType.FIXED_LARGE):
            raise McmError("implemented parameterisations: epsilon / fixed_small (configs/mcm/*) and start_x / fixed_large "
                           "(configs/stmogen/*); previous_x and learned variances are not used by the reference configs")
        if rescale_timesteps:
            raise McmError("rescale_timesteps is not used by the reference configs")
        self.opt = opt
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.timestep_map = list(range(self.num_timesteps))

    # ------------------------------------------------------------------ helpers
    def _tables(self):
        t = {k: getattr(self, k) for k in SamplerTables.FIELDS}
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            # p_mean_variance :527-531: "for fixedlarge, we set the initial (log-)variance like so to get a better decoder
            # log likelihood" -- only the DDPM update reads it (DDIM with eta = 0 uses no variance)
            t["posterior_log_variance_clipped"] = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
        return t

    @property
    def _mean_name(self):
        return "start_x" if self.model_mean_type == ModelMeanType.START_X else "epsilon"

    @staticmethod
    def _check_unsupported(model_kwargs, cond_fn, denoised_fn, pre_seq, transl_req, clip_denoised):
        if clip_denoised:
            raise McmError("clip_denoised=True is never used by MotionDiffusion (diffusion_architecture.py:179,186)")
        if cond_fn is not None or denoised_fn is not None or pre_seq is not None or transl_req is not None:
            raise McmError("cond_fn / denoised_fn / pre_seq / transl_req are not part of the re-hosted hot path")

    @staticmethod
    def _draw_seed(dev):
        """One draw from torch's generator of the sampling device: `torch.manual_seed` keeps controlling the run, as it
        controls the reference's randn_like draws (the streams themselves differ: Philox per step inside the library)."""
        return int(torch.randint(0, 2 ** 62, (1,), device=dev).item())

    def _draw_seed_async(self, salt=0):
        """A seed for the on-device noise WITHOUT a device->host read (the long-form pipeline must not synchronise):
        torch's CPU generator, which `torch.manual_seed` also controls."""
        return int(torch.randint(0, 2 ** 62, (1,)).item()) ^ (int(salt) * 0x9E3779B97F4A7C15 & (2 ** 62 - 1))

    def _run(self, model, shape, noise, model_kwargs, mode, eta, step_noise, device):
        B = shape[0]
        model_kwargs = dict(model_kwargs or {})
        dev = device if device is not None else next(model.parameters()).device
        if noise is None:
            noise = torch.randn(*shape, device=dev)
        # The reference draws randn_like(x) once per step (:685, :847).  Here the library generates step i's noise on the
        # device into ONE step-sized buffer (Philox keyed by (seed, i)); nothing of size n_steps x shape is allocated.
        # An explicit `step_noise` [n_steps, *shape] (device or host tensor) reproduces scripted reference draws.
        stochastic = (mode == "ddpm" and self.num_timesteps > 1) or (mode == "ddim" and eta != 0.0)
        seed = self._draw_seed(dev) if (stochastic and step_noise is None) else 0
        eng = model.bind_for_sampling(B, model_kwargs, dev)
        tables = SamplerTables(self._tables(), self.timestep_map, mode, eta, seed=seed, model_mean=self._mean_name)
        if step_noise is not None and step_noise.device.type == "cpu" and step_noise.numel() * 4 > (1 << 30):
            # large scripted noise stays on the host and is streamed one step at a time
            x0 = eng.sample_host(tables, noise.detach().float().cpu().contiguous(), None,
                                 step_noise.detach().float().contiguous())
            return x0.to(dev)
        return eng.sample(tables, noise.to(dev), step_noise)

    # ------------------------------------------------------------------ reference-facing API
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, pre_seq=None, transl_req=None, progress=False,
                      step_noise=None):
        """gaussian_diffusion.py:698-745.  `step_noise` [n_steps, *shape] (extension): the per-step noise the
        reference would draw; supply it for reproducible / oracle-comparable runs."""
        self._check_unsupported(model_kwargs, cond_fn, denoised_fn, pre_seq, transl_req, clip_denoised)
        return self._run(model, shape, noise, model_kwargs, "ddpm", 0.0, step_noise, device)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, pre_seq=None, step_noise=None,
                         repaint_noise=None):
        """gaussian_diffusion.py:925-997."""
        self._check_unsupported(model_kwargs, cond_fn, denoised_fn, pre_seq, None, clip_denoised)
        if self.opt is not None and getattr(self.opt, "same_overlap_noisy", False):
            raise McmError("opt.same_overlap_noisy (long-form RePaint) is a 'next' row, not implemented yet")
        y = (model_kwargs or {}).get("y") or {}
        mask = y.get("outpainting_mask", None) if hasattr(y, "get") else None
        if mask is not None and bool(torch.as_tensor(mask).any()):
            return self._run_repaint(model, shape, noise, model_kwargs, float(eta), device, repaint_noise)
        return self._run(model, shape, noise, model_kwargs, "ddim", float(eta), step_noise, device)

    def _run_repaint(self, model, shape, noise, model_kwargs, eta, device, repaint_noise):
        """ddim_sample_loop with y['outpainting_mask'] set (gaussian_diffusion.py:962-989): the harmonising loop unless
        opt.no_repaint, the mask blend of ddim_sample in every step.  `repaint_noise` [n_draws, *shape] (extension)
        replaces the reference's randn_like draws in order (scheduler.count_draws)."""
        opt = self.opt
        if opt is None:
            raise McmError("outpainting needs the `opt` namespace the reference's tools pass (overlap_len, addBlend, ...)")
        if eta != 0.0:
            raise McmError("outpainting is implemented for eta = 0 (what MotionDiffusion passes)")
        y = model_kwargs["y"]
        if "gt" not in y:
            raise McmError("y['outpainting_mask'] without y['gt']")
        times = None
        if not getattr(opt, "no_repaint", False):
            n = int(str(opt.timestep_respacing)[4:])                                      # 'ddim50' -> 50 (:1079-1084)
            if getattr(opt, "no_resample", False):
                times = get_schedule_jump_cjm_ddim(n)
            else:
                times = get_schedule_jump_cjm_ddim(n, jump_length=opt.jump_length, jump_n_sample=opt.jump_n_sample)
            if max(times) >= self.num_timesteps:
                raise McmError(f"harmonising schedule reaches step {max(times)} but the sampler has {self.num_timesteps}")
        B = shape[0]
        dev = device if device is not None else next(model.parameters()).device
        if noise is None:
            noise = torch.randn(*shape, device=dev)
        n_draws = count_draws(times, self.num_timesteps)
        seed = 0
        if repaint_noise is None:
            seed = self._draw_seed(dev)        # draws are generated inside the library, one buffer (mcm_repaint.noise_seq = NULL)
        elif repaint_noise.shape[0] < n_draws:
            raise McmError(f"repaint_noise holds {repaint_noise.shape[0]} draws, the schedule needs {n_draws}")
        eng = model.bind_for_sampling(B, dict(model_kwargs), dev)
        tables = SamplerTables(self._tables(), self.timestep_map, "ddim", 0.0, seed=seed, model_mean=self._mean_name)
        return eng.sample_repaint(tables, noise.to(dev), torch.as_tensor(y["gt"]), torch.as_tensor(y["outpainting_mask"]),
                                  repaint_noise, times=times, betas=self.betas, overlap_len=int(getattr(opt, "overlap_len", 0)),
                                  add_blend=bool(getattr(opt, "addBlend", True)))

    def training_losses(self, *a, **k):
        raise McmError("training is out of scope for motioncraft_b200 (inference hot path only)")


class SpacedDiffusion(GaussianDiffusion):
    """gaussian_diffusion.py:1407-1449: keep a subset of the base steps, re-derive betas."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        tmap, new_betas, last = [], [], 1.0
        for i, a in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - a / last)
                last = a
                tmap.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        self.timestep_map = tmap


def build_diffusion(cfg, opt=None):
    """mogen/models/architectures/diffusion_architecture.py:25-54."""
    betas = get_named_beta_schedule(cfg["beta_scheduler"], cfg["diffusion_steps"])
    mean = {"start_x": ModelMeanType.START_X, "previous_x": ModelMeanType.PREVIOUS_X,
            "epsilon": ModelMeanType.EPSILON}[cfg["model_mean_type"]]
    var = {"learned": ModelVarType.LEARNED, "fixed_small": ModelVarType.FIXED_SMALL,
           "fixed_large": ModelVarType.FIXED_LARGE, "learned_range": ModelVarType.LEARNED_RANGE}[cfg["model_var_type"]]
    if cfg.get("respace", None) is not None:
        return SpacedDiffusion(use_timesteps=space_timesteps(cfg["diffusion_steps"], cfg["respace"]), betas=betas,
                               model_mean_type=mean, model_var_type=var, loss_type=LossType.MSE, opt=opt)
    return GaussianDiffusion(betas=betas, model_mean_type=mean, model_var_type=var, loss_type=LossType.MSE, opt=opt)
