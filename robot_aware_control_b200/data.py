"""Training-data path: what happens between the stored clip and `SVGTrainer.train_step`.

Reference: each loader worker turns a clip's uint8 frames into float tensors on CPU -- ToTensor, and for the training
split a random crop + bilinear resize back to H x W and a shuffled colour jitter, masks cast back to {0, 1}
(src/dataset/robonet/robonet_dataset.py:257-300, 546-573) -- the DataLoader collates them batch-first, and
`process_batch` (:434-451) transposes every tensor to time-first and copies it to the device.

Here the loader only has to hand over the raw clip (uint8 frames (B, T, H, W, 3) and masks (B, T, H, W)): a quarter of
the host->device bytes, and one CUDA launch (`rac_process_batch`, csrc/data_kernels.cu) does ToTensor, crop / resize,
colour jitter, mask binarisation and the time-first layout. The random draws stay on the host: `sample_augment` consumes
python's `random` and torch's default generator exactly as the reference does, so a seeded run picks the same crop
window, factors and transform order.

Stored frames of another size than 48 x 64 (RoboNet stores 240 x 320) go through the dataset's `tf.Resize((48, 64))`
first, inside the same launch (bilinear without antialiasing: the torchvision 0.8 / 0.9 the reference pins).

The per-clip state / action glue of the dataset class (`_load_bounds`, `_load_states`, `_load_actions` with the
autograsp column, `_preprocess_bounds`, `_preprocess_states`, `_preprocess_actions`; robonet_dataset.py:173-255,
302-393) is `preprocess_states_actions`: the bounds / calibration records are assembled on the host (a handful of
numbers per clip, `clip_calibration`), the normalisation, camera-frame transform, action imputation and time-first
layout of the whole batch are one launch (`rac_preprocess_states`).

Not built: reading HDF5 itself (h5py is not in this image; any loader that yields the stored arrays works).
"""
import random

import numpy as np
import torch

from . import _lib

BRIGHTNESS, CONTRAST, SATURATION, HUE = 0, 1, 2, 3
TRANSPOSE_KEYS = ("qpos", "images", "states", "actions", "masks", "heatmaps", "raw_actions", "raw_states")

AUGMENT_DTYPE = np.dtype([("crop", np.int32, 4), ("factor", np.float64, 4), ("order", np.int32, 4)], align=True)
assert AUGMENT_DTYPE.itemsize == 64  # rac_augment (include/racb200.h)


def sample_augment(image_height=48, image_width=64):
    """One clip's augmentation parameters, drawn like robonet_dataset.py:261-275 + get_random_color_jitter (:546-573):
    random.randint(0, 5) -> crop size; RandomCrop.get_params (two torch.randint draws unless the crop is the full frame);
    four random.uniform factors (brightness, contrast, saturation in [0.8, 1.2], hue in [-0.1, 0.1]); random.shuffle of
    the four transforms. Returns (i, j, th, tw, [factors], [order])."""
    r = random.randint(0, 5)
    th, tw = image_height - r, image_width - r
    if (th, tw) == (image_height, image_width):
        i = j = 0
    else:
        i = int(torch.randint(0, image_height - th + 1, size=(1,)).item())
        j = int(torch.randint(0, image_width - tw + 1, size=(1,)).item())
    factors = [random.uniform(1 - 0.2, 1 + 0.2), random.uniform(1 - 0.2, 1 + 0.2), random.uniform(1 - 0.2, 1 + 0.2),
               random.uniform(-0.1, 0.1)]
    order = [BRIGHTNESS, CONTRAST, SATURATION, HUE]
    random.shuffle(order)
    return i, j, th, tw, factors, order


def pack_augment(augs, image_height=48, image_width=64):
    """List of per-clip tuples (as sample_augment returns) -> numpy structured array of rac_augment."""
    out = np.zeros(len(augs), AUGMENT_DTYPE)
    for b, (i, j, th, tw, factors, order) in enumerate(augs):
        if not (0 <= i and 0 <= j and 1 <= th and 1 <= tw and i + th <= image_height and j + tw <= image_width):
            raise ValueError(f"crop window {(i, j, th, tw)} outside the {image_height} x {image_width} frame")
        order = list(order) + [-1] * (4 - len(order))
        if len(factors) != 4 or len(order) != 4 or any(o not in (-1, 0, 1, 2, 3) for o in order):
            raise ValueError("augmentation needs 4 factors and at most 4 transform codes in 0..3")
        if not -0.5 <= factors[HUE] <= 0.5:
            raise ValueError(f"hue_factor ({factors[HUE]}) is not in [-0.5, 0.5].")
        if factors[CONTRAST] < 0 or factors[BRIGHTNESS] < 0 or factors[SATURATION] < 0:
            raise ValueError("brightness / contrast / saturation factors must be non-negative")
        out[b]["crop"] = (i, j, th, tw)
        out[b]["factor"] = factors
        out[b]["order"] = order
    return out


def preprocess_clips(frames, masks=None, augment=None, device=None):
    """uint8 frames (B, T, H, W, 3) [+ masks (B, T, H, W) float32 / uint8 / bool] -> time-first float32 device tensors
    images (T, B, 3, H, W), masks (T, B, 1, H, W). `augment`: None (evaluation splits: ToTensor only) or one
    sample_augment() tuple per clip."""
    lib = _lib.load()
    device = torch.device("cuda") if device is None else torch.device(device)
    frames = torch.as_tensor(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3:
        raise ValueError("frames must be uint8 (B, T, H, W, 3)")
    B, T, H, W = (int(s) for s in frames.shape[:4])  # H x W = the stored size; the model's frames are 48 x 64
    OH, OW = 48, 64
    frames = frames.to(device, non_blocking=True).contiguous()
    images = torch.empty(T, B, 3, OH, OW, device=device, dtype=torch.float32)
    m_dev = m_out = None
    mask_u8 = 0
    if masks is not None:
        masks = torch.as_tensor(masks)
        if tuple(masks.shape) != (B, T, H, W):
            raise ValueError(f"masks must be (B, T, H, W) = {(B, T, H, W)}, got {tuple(masks.shape)}")
        if masks.dtype == torch.bool:
            masks = masks.to(torch.uint8)
        if masks.dtype not in (torch.uint8, torch.float32):
            masks = masks.to(torch.float32)
        mask_u8 = int(masks.dtype == torch.uint8)
        m_dev = masks.to(device, non_blocking=True).contiguous()
        m_out = torch.empty(T, B, 1, OH, OW, device=device, dtype=torch.float32)
    aug_dev = None
    if augment is not None:
        if len(augment) != B:
            raise ValueError(f"{len(augment)} augmentation tuples for {B} clips")
        aug_dev = torch.from_numpy(pack_augment(augment, OH, OW).view(np.uint8).reshape(-1)).to(device)
    if B * T:
        _lib.check(lib.rac_process_batch(_lib.ptr(frames), _lib.ptr(m_dev) if m_dev is not None else None, mask_u8, B, T,
                                         H, W, _lib.ptr(aug_dev) if aug_dev is not None else None, _lib.ptr(images),
                                         _lib.ptr(m_out) if m_out is not None else None, _lib.stream_ptr()),
                   None, "rac_process_batch")
    return images, m_out


def process_batch(data, device, augment=None):
    """Drop-in for robonet_dataset.process_batch (:434-451): every tensor of the collated batch becomes time-first on
    `device`. Additionally, when data["images"] holds raw uint8 frames (B, T, H, W, 3) the per-clip preprocessing of
    the dataset class runs here on the device (preprocess_clips); float images (already preprocessed by a reference
    loader) are only transposed, as in the reference."""
    out = dict(data)
    imgs = data.get("images")
    raw = imgs is not None and torch.as_tensor(imgs).dtype == torch.uint8
    if raw:
        out["images"], m = preprocess_clips(imgs, data.get("masks"), augment, device)
        if m is not None:
            out["masks"] = m
    elif augment is not None:
        raise ValueError("augmentation on the device needs the raw uint8 frames")
    for k in TRANSPOSE_KEYS:
        if k in data and not (raw and k in ("images", "masks")):
            out[k] = torch.as_tensor(data[k]).transpose(1, 0).to(device, non_blocking=True)
    return out


# ---------------------------------------------------------------------------------------------- states / actions
LOCO_FRANKA_DIFF = (-0.365, -0.06103333)  # robonet_dataset.py:22
CLIP_CALIB_DTYPE = np.dtype([("kind", np.int32), ("camera", np.int32), ("grip_col", np.int32), ("pad_", np.int32),
                             ("low", np.float64, 5), ("high", np.float64, 5), ("world2cam", np.float64, 16),
                             ("frame_diff", np.float64, 2), ("grip_low", np.float64), ("grip_high", np.float64)], align=True)
assert CLIP_CALIB_DTYPE.itemsize == 256  # rac_clip_calib (include/racb200.h)


def load_bounds(robot_viewpoint, file_low=None, file_high=None):
    """RoboNetDataset._load_bounds (:196-206): fixed workspace box for the locobot / franka data, the file's
    low_bound / high_bound otherwise."""
    if "locobot" in robot_viewpoint or "franka" in robot_viewpoint:
        return (np.array([0.015, -0.3, 0.1, 0, 0], dtype=np.float32), np.array([0.55, 0.3, 0.4, 1, 1], dtype=np.float32))
    if file_low is None or file_high is None:
        raise ValueError(f"{robot_viewpoint}: RoboNet files carry their bounds (low_bound / high_bound)")
    return np.asarray(file_low), np.asarray(file_high)


def preprocess_bounds(low, high, preprocess_action="raw", world2cam=None):
    """RoboNetDataset._preprocess_bounds (:222-255): with a camera-frame action space the workspace box is projected
    into the camera frame and its axis-aligned hull becomes the normalisation box."""
    low, high = np.array(low, copy=True), np.array(high, copy=True)
    if "camera" in preprocess_action:
        if world2cam is None:
            raise ValueError("a camera-frame preprocess_action needs the world-to-camera matrix of the viewpoint")
        corners = np.array([[x, y, z, 1.0] for x in (low[0], high[0]) for y in (low[1], high[1]) for z in (low[2], high[2])]).T
        cam = (np.asarray(world2cam, np.float64) @ corners).T[:, :3]
        low[:3] = cam.min(0)
        high[:3] = cam.max(0)
    return low, high


def clip_calibration(robot_viewpoint, preprocess_action="raw", file_low=None, file_high=None, world2cam=None,
                     stored_state_dim=5):
    """One rac_clip_calib record (numpy, CLIP_CALIB_DTYPE) for a clip of `robot_viewpoint`: what _load_bounds and
    _preprocess_bounds compute per __getitem__ (:104-121)."""
    raw_low, raw_high = load_bounds(robot_viewpoint, file_low, file_high)
    low, high = preprocess_bounds(raw_low, raw_high, preprocess_action, world2cam)
    rec = np.zeros((), CLIP_CALIB_DTYPE)
    rec["kind"] = 1 if "locobot" in robot_viewpoint else (2 if "franka" in robot_viewpoint else 0)
    rec["camera"] = int("camera" in preprocess_action)
    rec["grip_col"] = stored_state_dim - 1
    rec["low"], rec["high"] = np.asarray(low, np.float64)[:5], np.asarray(high, np.float64)[:5]
    rec["world2cam"] = (np.eye(4) if world2cam is None else np.asarray(world2cam, np.float64)).reshape(16)
    rec["frame_diff"] = LOCO_FRANKA_DIFF
    rec["grip_low"], rec["grip_high"] = float(raw_low[4]), float(raw_high[4])
    return rec


def preprocess_states_actions(states, actions, calibs, robot_dim=5, action_dim=None, preprocess_action="raw",
                              impute_autograsp_action=True, device=None):
    """Collated stored states (B, T, S) and actions (B, T-1, A) -> what the reference dataset + process_batch hand to
    the trainer: normalised (camera-frame) states (T, B, robot_dim) and actions (T-1, B, action_dim), float32 on the
    device, time-first. `calibs`: one clip_calibration() record per clip. Raises like the reference for an action
    width that cannot be reconciled (:193-194) or a preprocess_action it does not implement (:341-352)."""
    lib = _lib.load()
    device = torch.device("cuda") if device is None else torch.device(device)
    if preprocess_action not in ("raw", "camera_raw"):
        raise NotImplementedError(preprocess_action)
    states = torch.as_tensor(np.asarray(states)).to(torch.float32)
    B, T, S = (int(v) for v in states.shape)
    if len(calibs) != B:
        raise ValueError(f"{len(calibs)} calibration records for {B} clips")
    if S > robot_dim:
        raise AssertionError("stored states are wider than cfg.robot_dim")  # (the reference asserts, :211)
    cal = np.stack([np.asarray(c) for c in calibs]).astype(CLIP_CALIB_DTYPE)
    if np.any(cal["grip_col"] >= S):
        raise ValueError("grip_col beyond the stored state width")
    if S < robot_dim:  # _load_states pads with zero columns (:210-213)
        states = torch.nn.functional.pad(states, (0, robot_dim - S))
    states = states.to(device).contiguous()
    a_dev = a_out = None
    A_in = A_out = 0
    if actions is not None:
        actions = torch.as_tensor(np.asarray(actions)).to(torch.float32)
        A_in = int(actions.shape[-1])
        A_out = A_in if action_dim is None else int(action_dim)
        if tuple(actions.shape[:2]) != (B, T - 1):
            raise ValueError(f"actions must be (B, T-1, A) = {(B, T - 1)}, got {tuple(actions.shape)}")
        if not (A_out == A_in or (impute_autograsp_action and A_out == A_in + 1)):
            raise ValueError(f"file adim {A_in}, target adim {A_out}")
        a_dev = actions.to(device).contiguous()
        a_out = torch.empty(T - 1, B, A_out, device=device, dtype=torch.float32)
    s_out = torch.empty(T, B, robot_dim, device=device, dtype=torch.float32)
    cal_dev = torch.from_numpy(cal.view(np.uint8).reshape(-1)).to(device)
    _lib.check(lib.rac_preprocess_states(_lib.ptr(states), _lib.ptr(a_dev), _lib.ptr(cal_dev), B, T, robot_dim, A_in, A_out,
                                         _lib.ptr(s_out), _lib.ptr(a_out), _lib.stream_ptr()), None, "rac_preprocess_states")
    if a_out is not None and preprocess_action == "camera_raw":
        # _make_camera_actions replaces the recorded actions by zeros before it uses them (`np.zeros_like`, :375): the
        # camera-frame displacement it returns is identically zero in every column. Reproduced, not repaired.
        a_out.zero_()
    return s_out, a_out
