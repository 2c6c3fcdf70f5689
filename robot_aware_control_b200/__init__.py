"""robot_aware_control_b200 -- B200-native (sm_100a) implementation of the planning hot path of
penn-pal-lab/robot_aware_control: batched SVG video-prediction rollouts inside the CEM policy plus the robot-aware
planning cost, behind the reference's own `SVGConvModel` (src/prediction) and `CEMPolicy` / `TrajectorySampler`
(src/cem) interfaces. All arithmetic runs in libracb200.so (hand-written CUDA, C ABI in include/racb200.h); there is
no CPU or library fallback -- importing the compute classes without the built extension raises.
"""
from .config import svg_config_from  # noqa: F401
from .state import State, DemoGoalState  # noqa: F401
from .model import SVGConvModel  # noqa: F401
from .losses import RobotWorldCost, ImgL2Cost, ImgDontcareCost, RobotL2Cost  # noqa: F401
from .losses import l1_criterion, dontcare_l1_criterion, kl_criterion  # noqa: F401
from .losses import robot_mse_criterion, world_mse_criterion, mse_criterion, dontcare_mse_criterion  # noqa: F401
from .robot import DeviceRobotModel  # noqa: F401
from .metrics import psnr, ssim, world_psnr_criterion  # noqa: F401
from .image import zero_robot_region  # noqa: F401
from .cem import CEMPolicy, TrajectorySampler  # noqa: F401
from .trainer import SVGTrainer  # noqa: F401
from .data import process_batch, preprocess_clips, sample_augment  # noqa: F401
from .data import clip_calibration, preprocess_bounds, preprocess_states_actions  # noqa: F401

__all__ = [
    "SVGConvModel", "CEMPolicy", "TrajectorySampler", "RobotWorldCost", "ImgL2Cost", "ImgDontcareCost",
    "RobotL2Cost", "State", "DemoGoalState", "zero_robot_region", "l1_criterion", "dontcare_l1_criterion",
    "kl_criterion", "robot_mse_criterion", "world_mse_criterion", "svg_config_from", "SVGTrainer",
    "psnr", "ssim", "world_psnr_criterion", "process_batch", "preprocess_clips", "sample_augment",
    "mse_criterion", "dontcare_mse_criterion", "DeviceRobotModel", "clip_calibration", "preprocess_bounds",
    "preprocess_states_actions",
]
