"""MotionDiffusion -- the architecture wrapper `tools/*.py` build through `build_architecture(cfg.model)`.

Mirrors the EVAL branch of mogen/models/architectures/diffusion_architecture.py:57-204 and
`BaseArchitecture.split_results` (base_architecture.py:112-140).  Training (`training_losses`, loss
parsing) is out of scope and raises.  Unlike the reference it does not construct the unused
`SMPLX_Skeleton(device="cuda")` (diffusion_architecture.py:80) which needs a hard-coded .npy file.
"""
import torch
from torch import nn

from ._lib import McmError
from .diffusion import build_diffusion
from .registry import ARCHITECTURES, LOSSES, build_submodule


@LOSSES.register_module()
class MSELoss(nn.Module):
    """Registered so `loss_recon=dict(type='MSELoss', ...)` in configs/mcm/* resolves; training-only."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, *a, **k):
        raise McmError("training losses are out of scope for motioncraft_b200")


def _to_cpu(v):
    return v.detach().cpu() if isinstance(v, torch.Tensor) else v


@ARCHITECTURES.register_module()
class MotionDiffusion(nn.Module):
    def __init__(self, model=None, loss_recon=None, loss_reduction="frame", diffusion_train=None,
                 diffusion_test=None, sampler_type="uniform", init_cfg=None, inference_type="ddpm", opt=None,
                 hand_loss_factor=1.0, face_no_loss=False, hand_no_loss=False, **kwargs):
        super().__init__()
        self.init_cfg = init_cfg
        self.inference_type = inference_type
        self.loss_reduction = loss_reduction
        self.hand_loss_factor, self.face_no_loss, self.hand_no_loss = hand_loss_factor, face_no_loss, hand_no_loss
        if inference_type != "gt":
            self.model = build_submodule(model) if isinstance(model, dict) or hasattr(model, "items") else model
        self.loss_recon = LOSSES.build(loss_recon) if loss_recon is not None else None
        self.diffusion_train = build_diffusion(diffusion_train) if diffusion_train is not None else None
        self.diffusion_test = build_diffusion(diffusion_test, opt=opt)
        self.eval()

    def forward(self, **kwargs):
        """diffusion_architecture.py:91-104 + :163-204."""
        if self.training:
            raise McmError("motioncraft_b200 implements the inference path only: call .eval() first")
        if not kwargs.get("return_loss", False) is False:
            raise McmError("return_loss=True (training) is out of scope")
        kwargs.pop("return_loss", None)
        motion = kwargs["motion"].float()
        motion_mask = kwargs["motion_mask"].float()
        B, T = motion.shape[:2]
        text = [kwargs["motion_metas"][i]["text"] for i in range(B)] if "motion_metas" in kwargs else None
        dim_pose = motion.shape[-1]
        if self.inference_type == "gt":
            output = motion
        else:
            dev = next(self.model.parameters()).device
            mk = self.model.get_precompute_condition(device=dev, text=text, **kwargs)
            mk["motion_mask"] = motion_mask
            mk["sample_idx"] = kwargs.get("sample_idx", None)
            mk["motion_length"] = kwargs["motion_length"]
            mk["num_intervals"] = kwargs.get("num_intervals", 1)
            mk["c"] = kwargs.get("c", None)
            mk["y"] = kwargs.get("y", {})
            mk["patch_size"] = kwargs.get("patch_size", 1)
            inference_kwargs = kwargs.get("inference_kwargs", {}) or {}
            if self.inference_type == "ddpm":
                output = self.diffusion_test.p_sample_loop(self.model, (B, T, dim_pose), clip_denoised=False,
                                                           progress=False, model_kwargs=mk, **inference_kwargs)
            elif self.inference_type == "ddim":
                output = self.diffusion_test.ddim_sample_loop(self.model, (B, T, dim_pose), clip_denoised=False,
                                                              progress=False, model_kwargs=mk, eta=0,
                                                              **inference_kwargs)
            else:
                raise McmError(f"unknown inference_type {self.inference_type!r}")
            if getattr(self.model, "post_process", None) is not None:
                output = self.model.post_process(output)
        results = kwargs
        results["pred_motion"] = output
        return self.split_results(results)

    @staticmethod
    def split_results(results):
        """base_architecture.py:112-140, with ONE device->host copy per tensor instead of one per sample."""
        B = results["motion"].shape[0]
        host = {k: _to_cpu(results[k]) for k in ("motion", "pred_motion", "motion_length", "motion_mask",
                                                 "pred_motion_length", "pred_motion_mask") if k in results}
        out = []
        for i in range(B):
            item = dict(motion=host["motion"][i], pred_motion=host["pred_motion"][i],
                        motion_length=host["motion_length"][i], motion_mask=host["motion_mask"][i])
            item["pred_motion_length"] = host.get("pred_motion_length", host["motion_length"])[i]
            item["pred_motion_mask"] = host.get("pred_motion_mask", host["motion_mask"])[i]
            if "motion_metas" in results:
                metas = results["motion_metas"][i]
                if "text" in metas:
                    item["text"] = metas["text"]
                if "token" in metas:
                    item["token"] = metas["token"]
            out.append(item)
        return out
