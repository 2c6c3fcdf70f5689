"""TEST INFRASTRUCTURE -- CPU restatement (numpy, the reference's own dtype flow) of the per-clip state / action glue
of the reference dataset class, src/dataset/robonet/robonet_dataset.py:
  _load_bounds :196-206, _load_states :208-214, _load_actions :173-194 (autograsp imputation), _preprocess_bounds
  :222-255 (workspace box projected into the camera frame), _preprocess_states :302-334, _preprocess_actions /
  _make_camera_actions :336-393, normalize / denormalize :470-480.
Pinned by tests/golden/dataset_glue.npz, which oracle/make_golden_dataset.py generates by calling those UNMODIFIED
reference methods. Only tests/ may import this module."""
import numpy as np

LOCO_FRANKA_DIFF = np.array([-0.365, -0.06103333])  # robonet_dataset.py:22


def denormalize(states, low, high):
    return states * (high - low) + low


def normalize(states, low, high):
    return (states - low) / (high - low)


def load_bounds(robot_viewpoint, file_low=None, file_high=None):
    if "locobot" in robot_viewpoint or "franka" in robot_viewpoint:
        return (np.array([0.015, -0.3, 0.1, 0, 0], dtype=np.float32), np.array([0.55, 0.3, 0.4, 1, 1], dtype=np.float32))
    return file_low, file_high


def load_states(file_states, robot_dim):
    states = file_states.astype(np.float32)
    if states.shape[-1] != robot_dim:
        assert robot_dim > states.shape[-1]
        states = np.pad(states, [(0, 0), (0, robot_dim - states.shape[-1])])
    return states


def load_actions(file_actions, file_states, gripper_low, gripper_high, action_dim, impute_autograsp):
    actions = file_actions.astype(np.float32)
    a_T, adim = actions.shape
    if action_dim == adim:
        return actions
    if impute_autograsp and adim + 1 == action_dim:
        nxt = file_states[1:, -1]
        # (the reference indexes gripper_high[-1] of the scalar raw_high[4] it was given: a 0-d value, see the generator)
        mid = (gripper_high + gripper_low) / 2.0
        col = np.where(nxt > mid, gripper_high, gripper_low).reshape(a_T, 1)
        return np.concatenate((actions, col), axis=-1).astype(np.float32)
    raise ValueError(f"file adim {adim}, target adim {action_dim}")


def preprocess_bounds(low, high, preprocess_action, world2cam=None):
    low, high = low.copy(), high.copy()
    if "camera" in preprocess_action:
        xs, ys, zs = (low[0], high[0]), (low[1], high[1]), (low[2], high[2])
        box = np.array([[x, y, z] for x in xs for y in ys for z in zs])
        box = np.concatenate([box, np.ones((8, 1))], 1).T
        cbox = ((world2cam @ box).T)[:, :3]
        low[:3] = np.min(cbox, 0)
        high[:3] = np.max(cbox, 0)
    return low, high


def preprocess_states(states, low, high, robot_viewpoint, preprocess_action, world2cam=None):
    states = states.copy()
    if "locobot" in robot_viewpoint:
        eef = states[:, :3]
    elif "franka" in robot_viewpoint:
        eef = states[:, :3]
        eef[:, :2] += LOCO_FRANKA_DIFF
        eef[:, 2] = 0.14
    else:
        eef = denormalize(states[:, :3], low[:3], high[:3])
    if "camera" in preprocess_action:
        eef = np.concatenate([eef, np.ones((eef.shape[0], 1))], 1).T
        eef = ((world2cam @ eef).T)[:, :3]
    states[:, :3] = normalize(eef, low[:3], high[:3])
    states[:, 4] = normalize(states[:, 4], low[4], high[4])
    return states


def preprocess_actions(states, actions, preprocess_action):
    """"raw": the recorded actions. "camera_raw": _make_camera_actions REPLACES the recorded actions by zeros before it
    uses them (`actions = np.zeros_like(actions)`, :375), so the camera-frame displacement of (s, s + a) it returns is
    identically zero in every column -- restated as such."""
    if preprocess_action == "raw":
        return actions
    if preprocess_action == "camera_raw":
        return np.zeros_like(actions)
    raise NotImplementedError(preprocess_action)
