"""TEST INFRASTRUCTURE -- golden vectors of the rest of the training-data path: the stored-size
`tf.Resize` in front of _preprocess_images_masks and the per-clip state / action glue of the UNMODIFIED reference
RoboNetDataset (src/dataset/robonet/robonet_dataset.py:173-255,257-300,302-393), run on CPU over synthetic clips.
    python -m oracle.make_golden_dataset      -> tests/golden/dataset_glue.npz

torchvision: the reference pins 0.8.1 / 0.9.1, where tf.Resize of a float tensor is plain bilinear interpolation
(antialiasing arrived in 0.10 and became the default for tensors in 0.17). 0.26 is installed here, so the dataset's
transform is built as Compose([ToTensor(), Resize((48, 64), antialias=False)]): the pinned behaviour."""
import importlib
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.make_golden_data import ORDER  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
H, W, T = 48, 64, 2


def synth_stored(rs, hs, ws):
    low = rs.randint(0, 256, (T, 6, 8, 3)).astype(np.float32)
    img = np.repeat(np.repeat(low, hs // 6, 1), ws // 8, 2) + rs.randint(-25, 26, (T, hs, ws, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    mask = np.zeros((T, hs, ws), np.float32)
    for t in range(T):
        y, x = rs.randint(0, hs // 2), rs.randint(0, ws // 2)
        mask[t, y:y + rs.randint(hs // 12, hs // 3), x:x + rs.randint(ws // 16, ws // 3)] = 1.0
    return img, mask


def main():
    ref_shim.import_reference()
    mod = importlib.import_module("src.dataset.robonet.robonet_dataset")
    calib = importlib.import_module("src.utils.camera_calibration")
    out = {}
    # ------------------------------------------------------------------ stored-size frames
    real_F = mod.F
    log = []

    def logged(name):
        fn = getattr(real_F, name)

        def wrapper(img, *a):
            log.append((name,) + tuple(a))
            return fn(img, *a)
        return wrapper

    proxy = types.SimpleNamespace(**{k: getattr(real_F, k) for k in dir(real_F) if not k.startswith("__")})
    for name in list(ORDER) + ["crop"]:
        setattr(proxy, name, logged(name))
    mod.F = proxy
    ds = mod.RoboNetDataset.__new__(mod.RoboNetDataset)
    ds._config = types.SimpleNamespace(image_width=W, image_height=H)
    ds._img_transform = mod.tf.Compose([mod.tf.ToTensor(), mod.tf.Resize((H, W), antialias=False)])
    rs = np.random.RandomState(11)
    for tag, (hs, ws), nclips in (("s96", (96, 128), 2), ("s240", (240, 320), 2)):
        frames, masks, plain_i, plain_m, aug_i, aug_m, params = [], [], [], [], [], [], []
        for b in range(nclips):
            img, mask = synth_stored(rs, hs, ws)
            frames.append(img)
            masks.append(mask)
            ds._augment_img = False
            v, m = ds._preprocess_images_masks(img, mask)
            plain_i.append(v)
            plain_m.append(m)
            random.seed(300 + b)
            torch.manual_seed(300 + b)
            ds._augment_img = True
            del log[:]
            v, m = ds._preprocess_images_masks(img, mask)
            aug_i.append(v)
            aug_m.append(m)
            crop = [e for e in log if e[0] == "crop"][0][1:]
            first = [e for e in log if e[0] != "crop"][:4]
            factors = [0.0] * 4
            for name, f in first:
                factors[ORDER[name]] = f
            params.append(list(crop) + factors + [ORDER[name] for name, _ in first])
        out[f"{tag}_frames"] = np.stack(frames)
        out[f"{tag}_masks"] = np.stack(masks)
        out[f"{tag}_images_plain"] = torch.stack(plain_i).transpose(1, 0).contiguous().numpy()
        out[f"{tag}_masks_plain"] = torch.stack(plain_m).transpose(1, 0).contiguous().numpy()
        out[f"{tag}_images_aug"] = torch.stack(aug_i).transpose(1, 0).contiguous().numpy()
        out[f"{tag}_masks_aug"] = torch.stack(aug_m).transpose(1, 0).contiguous().numpy()
        out[f"{tag}_params"] = np.array(params, np.float64)
    mod.F = real_F
    # ------------------------------------------------------------------ states / actions / bounds
    TS = 6
    cases = [  # tag, robot viewpoint, preprocess_action, stored state dim, stored action dim, config action dim
        ("robonet_raw", "sawyer_sudri0_c0", "raw", 5, 4, 5),
        ("robonet_cam", "sawyer_sudri0_c1", "camera_raw", 5, 4, 4),
        ("locobot_raw", "locobot_c0", "raw", 5, 5, 5),
        ("locobot_cam", "locobot_modified_c0", "camera_raw", 4, 5, 5),
        ("franka_raw", "franka_c0", "raw", 5, 5, 5),
        ("franka_cam", "franka_c0", "camera_raw", 5, 4, 5),
    ]
    rs = np.random.RandomState(21)
    for tag, robot, pa, sdim, adim, cfg_adim in cases:
        ds = mod.RoboNetDataset.__new__(mod.RoboNetDataset)
        ds._config = types.SimpleNamespace(preprocess_action=pa, robot_dim=5, robot_joint_dim=6)
        ds._traj_robots = [robot]
        ds._action_dim = cfg_adim
        ds._impute_autograsp_action = True
        if "locobot" in robot or "franka" in robot:
            st = np.concatenate([rs.uniform([0.05, -0.25, 0.1], [0.5, 0.25, 0.4], (TS, 3)), rs.uniform(-1, 1, (TS, 1)),
                                 rs.uniform(0, 1, (TS, 1))], 1)[:, :sdim]
        else:
            st = np.concatenate([rs.uniform(0, 1, (TS, 3)), rs.uniform(-1, 1, (TS, 1)), rs.uniform(-1, 1, (TS, 1))], 1)[:, :sdim]
        fp = {"states": st.astype(np.float64), "actions": rs.uniform(-0.05, 0.05, (TS - 1, adim)),
              "low_bound": np.array([0.4, -0.3, 0.15, -1.5, -1.0]), "high_bound": np.array([0.85, 0.35, 0.45, 1.5, 1.0])}
        raw_low, raw_high = ds._load_bounds(fp, robot, 0)
        states = ds._load_states(fp, 0, TS)
        # the reference passes the scalar bounds raw_low[4] / raw_high[4] and then indexes [-1] (robonet_dataset.py:104,183):
        # that only works for array bounds, so the autograsp branch is exercised with 1-element arrays
        actions = ds._load_actions(fp, np.atleast_1d(raw_low[4]), np.atleast_1d(raw_high[4]), 0, TS - 1)
        low, high = ds._preprocess_bounds(raw_low, raw_high, 0)
        pstates = ds._preprocess_states(states, low, high, robot, 0)
        pactions = ds._preprocess_actions(pstates, actions, low, high, 0).numpy()
        for k, v in (("file_states", fp["states"]), ("file_actions", fp["actions"]), ("raw_low", raw_low), ("raw_high", raw_high),
                     ("loaded_states", states), ("loaded_actions", actions), ("low", low), ("high", high),
                     ("states", pstates), ("actions", pactions), ("world2cam", calib.world_to_camera_dict[robot])):
            out[f"{tag}_{k}"] = np.asarray(v)
        print(tag, pstates.dtype, pactions.dtype, np.abs(pactions).max())
    out["case_tags"] = np.array([c[0] for c in cases])
    out["case_robots"] = np.array([c[1] for c in cases])
    out["case_modes"] = np.array([c[2] for c in cases])
    out["case_action_dims"] = np.array([c[5] for c in cases])
    np.savez_compressed(os.path.join(OUT, "dataset_glue.npz"), **out)


if __name__ == "__main__":
    main()
