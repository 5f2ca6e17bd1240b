"""mogen/models/losses -- training-only; the registered name keeps `loss_recon=dict(type='MSELoss', ...)` resolvable."""
from motioncraft_b200.architecture import MSELoss  # noqa: F401

__all__ = ["MSELoss"]
