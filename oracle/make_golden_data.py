"""TEST INFRASTRUCTURE -- golden vectors of the training-data path: runs the UNMODIFIED
reference RoboNetDataset._preprocess_images_masks (src/dataset/robonet/robonet_dataset.py:257-300) and process_batch
(:434-451) on CPU over synthetic uint8 clips -- with and without augmentation -- and stores inputs, the augmentation
parameters the reference drew (captured by logging its torchvision calls) and the float outputs in
tests/golden/data_path.npz.   python -m oracle.make_golden_data

torchvision here is 0.26 (the reference pins 0.9.1): for the UP-scaling resize of the <= 5 px smaller crop, bilinear
with and without antialiasing are the same function, so the version difference is limited to float rounding."""
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

OUT = os.path.join(ROOT, "tests", "golden")
H, W, T, B = 48, 64, 2, 6
ORDER = {"adjust_brightness": 0, "adjust_contrast": 1, "adjust_saturation": 2, "adjust_hue": 3}


def synth_clip(rs):
    """Smooth-ish coloured frames (random low-res pattern, upsampled) + a rectangle mask, as uint8 / float32."""
    low = rs.randint(0, 256, (T, 6, 8, 3)).astype(np.float32)
    img = np.repeat(np.repeat(low, 8, 1), 8, 2) + rs.randint(-20, 21, (T, H, W, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    img[:, :4, :4] = 128  # a grey patch: max == min in rgb -> hsv (the guarded division)
    mask = np.zeros((T, H, W), np.float32)
    for t in range(T):
        y, x = rs.randint(0, H - 12), rs.randint(0, W - 12)
        mask[t, y:y + rs.randint(4, 12), x:x + rs.randint(4, 12)] = 1.0
    return img, mask


def main():
    ref_shim.import_reference()
    mod = importlib.import_module("src.dataset.robonet.robonet_dataset")
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
    ds._img_transform = mod.tf.Compose([mod.tf.ToTensor(), mod.tf.Resize((H, W))])
    rs = np.random.RandomState(5)
    out = {}
    frames, masks, imgs_plain, masks_plain, imgs_aug, masks_aug, params, seeds = [], [], [], [], [], [], [], []
    for b in range(B):
        img, mask = synth_clip(rs)
        frames.append(img)
        masks.append(mask)
        ds._augment_img = False
        v, m = ds._preprocess_images_masks(img, mask)
        imgs_plain.append(v)
        masks_plain.append(m)
        # augmentation: the reference draws from `random` and from torch's default generator (RandomCrop.get_params)
        seed = 100 + b
        random.seed(seed)
        torch.manual_seed(seed)
        seeds.append(seed)
        ds._augment_img = True
        del log[:]
        v, m = ds._preprocess_images_masks(img, mask)
        imgs_aug.append(v)
        masks_aug.append(m)
        crop = [e for e in log if e[0] == "crop"][0][1:]
        first = [e for e in log if e[0] != "crop"][:4]  # the four colour transforms of the first frame, in order
        assert sorted(e[0] for e in first) == sorted(ORDER)
        factors = [0.0] * 4
        for name, f in first:
            factors[ORDER[name]] = f
        params.append(list(crop) + factors + [ORDER[name] for name, _ in first])
        print(b, params[-1])
    # the collated batch is batch-first (B, T, ...); process_batch makes it time-first
    data = {"images": torch.stack(imgs_aug), "masks": torch.stack(masks_aug)}
    data = mod.process_batch(data, torch.device("cpu"))
    assert data["images"].shape == (T, B, 3, H, W) and data["masks"].shape == (T, B, 1, H, W)
    out["frames"] = np.stack(frames)                       # (B, T, H, W, 3) uint8
    out["masks"] = np.stack(masks)                         # (B, T, H, W) float32 {0, 1}
    out["images_plain"] = torch.stack(imgs_plain).transpose(1, 0).contiguous().numpy()
    out["masks_plain"] = torch.stack(masks_plain).transpose(1, 0).contiguous().numpy()
    out["images_aug"] = data["images"].contiguous().numpy()
    out["masks_aug"] = data["masks"].contiguous().numpy()
    out["params"] = np.array(params, np.float64)           # per clip: i, j, th, tw, b, c, s, h factors, order[4]
    out["seeds"] = np.array(seeds)
    np.savez_compressed(os.path.join(OUT, "data_path.npz"), **out)


if __name__ == "__main__":
    main()
