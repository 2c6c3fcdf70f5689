"""Device-side stand-in for the reference's analytical robot models inside the planner.

`TrajectorySampler.generate_model_rollouts` asks `robot_model.predict_batch(data, thick=True)` for the candidates' future
robot states and masks every CEM iteration (reference src/cem/trajectory_sampler.py:86-109). In the reference that is a
per-candidate Python loop around an IK call and a MuJoCo segmentation render (src/dataset/wx250s/wx250s_model.py:
121-182), i.e. a host round trip (`actions.cpu()`) per iteration. `DeviceRobotModel` keeps both halves on the GPU:

* states: `rac_predict_states` -- the planar end-effector integration of `WX250sAnalyticalModel` / `FrankaAnalyticalModel`
  with the reference's float32 / float64 mix, bit-equal to the reference (tests/golden/robot_states.npz).
* masks: `rac_render_masks` -- a capsule model of the arm posed by closed-form planar IK and rasterised through a
  pinhole camera. The reference's masks come from MuJoCo meshes + the Interbotix IK solver, neither of which exists
  outside the authors' setup: this half is NOT pinned to the reference (its oracle is oracle/robot_oracle.py).
  Pass `mask_fn=` (states (T+1,N,5) CUDA -> masks (T+1,N,1,H,W) CUDA) to plug another device-side mask source in.

`CEMPolicy(..., robot_model=DeviceRobotModel(...))` then runs the whole robot-aware plan as ONE `rac_cem_plan` call.
The object also answers `predict_batch(data, thick)` like the reference models (returning CUDA tensors), so it can be
handed to any code written against that interface.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

LOCO_FRANKA_DIFF = (-0.365, -0.06103333)   # src/utils/camera_calibration.py:176
LOCO_WX250S_DIFF = (-0.13, -0.01)          # src/utils/camera_calibration.py:177
# camera_to_world_dict["wx250s_c0"] (src/utils/camera_calibration.py:158-161)
WX250S_CAM_TO_WORLD = np.array([[0.05598868, 0.80338198, -0.592826, 0.82155341],
                                [0.99834883, -0.0526833, 0.02289275, -0.018],
                                [-0.01284041, -0.59312888, -0.80500513, 0.58407623],
                                [0.0, 0.0, 0.0, 1.0]])
# the 320 x 240 logitech intrinsics of cam_intrinsics_dict (:169) scaled to the model's 64 x 48 frames
DEFAULT_INTRINSICS = np.array([[64.15, 0.0, 32.0], [0.0, 64.15, 24.0], [0.0, 0.0, 1.0]])


class DeviceRobotModel(object):
    KINDS = {"wx250s": 0, "franka": 1}

    def __init__(self, config=None, kind="wx250s", push_height=0.1, default_pitch=1.5, cam_ext=None, intrinsics=None,
                 low=(0.015, -0.3, 0.1, 0, 0), high=(0.55, 0.3, 0.4, 1, 1), shoulder_z=0.11, l_upper=0.255,
                 l_fore=0.25, l_wrist=0.17, radius=(0.06, 0.035, 0.03, 0.035), thick_extra=0.01, render_masks=True,
                 mask_fn=None, height=48, width=64):
        if kind not in self.KINDS:
            raise ValueError(f"robot kind {kind!r}: one of {sorted(self.KINDS)}")
        self._config = config
        self.kind = kind
        self.height, self.width = height, width
        self.render_masks = bool(render_masks) and mask_fn is None
        self.mask_fn = mask_fn
        self.thick_extra = float(thick_extra)
        m = _lib.RacRobotModel()
        m.kind = self.KINDS[kind]
        for i in range(5):
            m.low[i], m.high[i] = float(low[i]), float(high[i])
        diff = LOCO_WX250S_DIFF if kind == "wx250s" else LOCO_FRANKA_DIFF
        m.frame_diff[0], m.frame_diff[1] = diff
        m.push_height = float(push_height)
        cam_ext = WX250S_CAM_TO_WORLD if cam_ext is None else np.asarray(cam_ext, dtype=np.float64)
        K = DEFAULT_INTRINSICS if intrinsics is None else np.asarray(intrinsics, dtype=np.float64)
        R, c = cam_ext[:3, :3], cam_ext[:3, 3]          # camera -> robot frame
        minv = R @ np.linalg.inv(K)                     # pixel (u, v, 1) -> ray direction in the robot frame
        for i in range(3):
            m.cam_center[i] = float(c[i])
        for i in range(9):
            m.cam_minv[i] = float(minv.reshape(-1)[i])
        m.shoulder_z, m.l_upper, m.l_fore, m.l_wrist = float(shoulder_z), float(l_upper), float(l_fore), float(l_wrist)
        m.pitch = float(default_pitch)
        for i in range(4):
            m.radius[i] = float(radius[i])
        self.c_model = m
        self._lib = _lib.load()

    # ---- device entry points -------------------------------------------------------------------------------------
    def predict_states(self, start_state_norm, actions):
        """start_state_norm: (5,) normalised start state (CUDA); actions (N, L, A) CUDA fp32 -> states (L+1, N, 5)."""
        n, L, A = actions.shape
        states = torch.empty(L + 1, n, 5, device=actions.device)
        _lib.check(self._lib.rac_predict_states(C.byref(self.c_model), _lib.ptr(start_state_norm),
                                                _lib.ptr(actions.contiguous()), n, L, A, _lib.ptr(states), n * 5,
                                                _lib.stream_ptr()), None, "rac_predict_states")
        return states

    def render(self, states, thick=True):
        """states (T, N, 5) normalised CUDA -> masks (T, N, 1, H, W) float {0, 1}."""
        if self.mask_fn is not None:
            return self.mask_fn(states)
        T1, n, _ = states.shape
        masks = torch.empty(T1, n, 1, self.height, self.width, device=states.device)
        _lib.check(self._lib.rac_render_masks(C.byref(self.c_model), _lib.ptr(states.contiguous()), n * 5, n, T1 - 1,
                                              self.height, self.width, self.thick_extra if thick else 0.0,
                                              _lib.ptr(masks), n * self.height * self.width, _lib.stream_ptr()),
                   None, "rac_render_masks")
        return masks

    # ---- the reference robot-model interface (wx250s_model.py:121-182) ------------------------------------------------
    @torch.no_grad()
    def predict_batch(self, data, thick=False):
        """data = {"states" (T+1, N, 5) with row 0 = normalised start state, "actions" (T, N, A), "low", "high", "qpos"}
        as built at trajectory_sampler.py:93-106. All candidates share the start state there; per-candidate bounds other
        than this model's are not supported. Returns (states, masks) as CUDA tensors."""
        dev = torch.device("cuda", torch.cuda.current_device())
        st = data["states"]
        if not bool((st[0] == st[0, :1]).all()):
            raise NotImplementedError("DeviceRobotModel.predict_batch: candidates with different start states")
        start = st[0, 0].to(dev, dtype=torch.float32).contiguous()
        actions = data["actions"].to(dev, dtype=torch.float32).permute(1, 0, 2).contiguous()  # (N, T, A)
        states = self.predict_states(start, actions)
        return states, self.render(states, thick)


def normalized_start_state(start_state, cfg_experiment, low, high):
    """trajectory_sampler.py:93-99: shift the robot-frame start state into the loco frame, then normalise (float32)."""
    s = torch.tensor(np.asarray(start_state, dtype=np.float32))
    # float32 tensor + float64 numpy constant: the reference's sum is formed in double and stored back as float32
    if cfg_experiment == "control_franka":
        s[:2] = s[:2] + torch.from_numpy(np.array(LOCO_FRANKA_DIFF))
    elif cfg_experiment == "control_wx250s":
        s[:2] = s[:2] + torch.from_numpy(np.array(LOCO_WX250S_DIFF))
    return (s - low.reshape(-1)) / (high.reshape(-1) - low.reshape(-1))
