"""mogen/models/utils/scheduler.py:178-208 (the RePaint jump schedule used by the long-form sampler)."""
from motioncraft_b200.scheduler import get_schedule_jump_cjm_ddim  # noqa: F401
