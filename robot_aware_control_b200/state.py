"""Interface dataclasses of the planner (reference src/utils/state.py:4-19): same field names and defaults, so
objects built by reference callers (controllers, episode runners) can be passed unchanged."""
from dataclasses import dataclass
from typing import Any


@dataclass
class State:
    img: Any = None        # uint8 (H, W, 3) start image, or a tensor inside the cost
    state: Any = None      # robot end-effector state
    sim_state: Any = None
    mask: Any = None       # robot mask
    sim: Any = None
    qpos: Any = None       # joint positions for the analytical robot model


@dataclass
class DemoGoalState:
    imgs: Any = None       # list of uint8 (H, W, 3) goal images
    states: Any = None
    sim_states: Any = None
    masks: Any = None      # list of float32 (1, H, W) goal masks
    qposes: Any = None
