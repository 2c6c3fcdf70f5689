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

Not built: reading HDF5 (h5py is not in this image), state / action normalisation (numpy glue of the dataset class),
stored frames of another size than 48 x 64 (the reference's `tf.Resize` down-scaling before everything else).
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
    B, T, H, W = (int(s) for s in frames.shape[:4])
    if (H, W) != (48, 64):
        raise NotImplementedError("the device data path handles stored 48 x 64 frames")
    frames = frames.to(device, non_blocking=True).contiguous()
    images = torch.empty(T, B, 3, H, W, device=device, dtype=torch.float32)
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
        m_out = torch.empty(T, B, 1, H, W, device=device, dtype=torch.float32)
    aug_dev = None
    if augment is not None:
        if len(augment) != B:
            raise ValueError(f"{len(augment)} augmentation tuples for {B} clips")
        aug_dev = torch.from_numpy(pack_augment(augment, H, W).view(np.uint8).reshape(-1)).to(device)
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
