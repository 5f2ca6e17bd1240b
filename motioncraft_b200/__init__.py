"""motioncraft_b200 -- B200-native (sm_100a) denoising hot path of cure-lab/MotionCraft (configs/mcm/*)."""
__version__ = "0.1.0"
