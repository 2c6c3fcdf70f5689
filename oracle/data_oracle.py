"""TEST INFRASTRUCTURE -- CPU restatement of the reference's training-data path: uint8 frames -> float tensors
(ToTensor), optional augmentation (random crop + bilinear resize back to H x W, colour jitter), masks re-binarised,
batch-first -> time-first. Follows src/dataset/robonet/robonet_dataset.py:257-300 (_preprocess_images_masks),
:546-573 (get_random_color_jitter) and :434-451 (process_batch).

The pixel arithmetic lives in a third-party dependency, torchvision (requirements.txt pins 0.9.1; 0.26 is installed
here): transforms.functional.{to_tensor, crop, resize, adjust_brightness, adjust_contrast, adjust_saturation,
adjust_hue} on float tensors, restated below with plain torch ops:
  resize  = torch.nn.functional.interpolate(mode="bilinear", align_corners=False) (ATen upsample_bilinear2d): source
            coordinate max(scale * (dst + 0.5) - 0.5, 0), scale = in / out, two taps per axis
  blend   = (ratio * a + (1 - ratio) * b).clamp(0, 1); brightness: b = 0; contrast: b = mean(gray); saturation: b = gray
  gray    = 0.2989 r + 0.587 g + 0.114 b
  hue     = rgb -> hsv (Pillow's formulas), h = (h + factor) % 1, hsv -> rgb
Pinned by tests/golden/data_path.npz (oracle/make_golden_data.py runs the unmodified reference) and, for clips stored
at another size, tests/golden/dataset_glue.npz (oracle/make_golden_dataset.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module."""
import math

import numpy as np
import torch

BRIGHTNESS, CONTRAST, SATURATION, HUE = 0, 1, 2, 3


def to_tensor(frames_u8):
    """tf.ToTensor on each HWC uint8 frame (robonet_dataset.py:58,279,294): CHW float32, value / 255."""
    x = torch.from_numpy(np.ascontiguousarray(frames_u8)).permute(0, 3, 1, 2).contiguous()
    return x.to(torch.float32).div(255)


def _axis_taps(n_in, n_out, offset):
    """Bilinear source taps of one axis (align_corners False), in float32 like ATen: index0, index1, lambda1."""
    scale = np.float32(n_in) / np.float32(n_out)
    i0 = np.zeros(n_out, np.int64)
    i1 = np.zeros(n_out, np.int64)
    l1 = np.zeros(n_out, np.float32)
    for d in range(n_out):
        real = np.float32(scale * np.float32(d + 0.5) - np.float32(0.5))
        if real < 0:
            real = np.float32(0)
        k = int(math.floor(real))
        k = min(k, n_in - 1)
        i0[d] = offset + k
        i1[d] = offset + k + (1 if k < n_in - 1 else 0)
        l1[d] = min(max(np.float32(real - np.float32(k)), np.float32(0)), np.float32(1))
    return i0, i1, l1


def crop_resize(x, i, j, th, tw, H, W):
    """F.resize(F.crop(x, i, j, th, tw), (H, W)) for x (..., H, W) float32 (robonet_dataset.py:281-288)."""
    if (th, tw) == (H, W):
        return x[..., i:i + th, j:j + tw].clone()
    y0, y1, ly = _axis_taps(th, H, i)
    x0, x1, lx = _axis_taps(tw, W, j)
    ly = torch.from_numpy(ly).view(-1, 1)
    lx = torch.from_numpy(lx).view(1, -1)
    p00 = x[..., y0, :][..., :, x0]
    p01 = x[..., y0, :][..., :, x1]
    p10 = x[..., y1, :][..., :, x0]
    p11 = x[..., y1, :][..., :, x1]
    return (1 - ly) * ((1 - lx) * p00 + lx * p01) + ly * ((1 - lx) * p10 + lx * p11)


def _gray(x):
    return 0.2989 * x[..., 0, :, :] + 0.587 * x[..., 1, :, :] + 0.114 * x[..., 2, :, :]


def _blend(a, b, ratio):
    ratio = float(ratio)
    return (ratio * a + (1.0 - ratio) * b).clamp(0, 1)


def adjust_hue(x, factor):
    r, g, b = x[..., 0, :, :], x[..., 1, :, :], x[..., 2, :, :]
    maxc = torch.maximum(torch.maximum(r, g), b)
    minc = torch.minimum(torch.minimum(r, g), b)
    eqc = maxc == minc
    cr = maxc - minc
    ones = torch.ones_like(maxc)
    s = cr / torch.where(eqc, ones, maxc)
    div = torch.where(eqc, ones, cr)
    rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
    hr = (maxc == r) * (bc - gc)
    hg = ((maxc == g) & (maxc != r)) * (2.0 + rc - bc)
    hb = ((maxc != g) & (maxc != r)) * (4.0 + gc - rc)
    h = torch.fmod((hr + hg + hb) / 6.0 + 1.0, 1.0)
    h = (h + factor) % 1.0
    v = maxc
    i = torch.floor(h * 6.0)
    f = h * 6.0 - i
    i = i.to(torch.int32) % 6
    p = (v * (1.0 - s)).clamp(0, 1)
    q = (v * (1.0 - s * f)).clamp(0, 1)
    t = (v * (1.0 - s * (1.0 - f))).clamp(0, 1)
    sel = lambda opts: sum((i == k) * o for k, o in enumerate(opts))
    return torch.stack((sel((v, q, p, p, t, v)), sel((t, v, v, q, p, p)), sel((p, p, t, v, v, q))), dim=-3)


def color_jitter(x, factors, order):
    """The shuffled Compose of get_random_color_jitter (robonet_dataset.py:546-573) on one (3, H, W) frame."""
    for op in order:
        f = float(factors[op])
        if op == BRIGHTNESS:
            x = _blend(x, torch.zeros_like(x), f)
        elif op == CONTRAST:
            x = _blend(x, _gray(x).mean(), f)
        elif op == SATURATION:
            x = _blend(x, _gray(x).unsqueeze(-3), f)
        elif op == HUE:
            x = adjust_hue(x, f)
    return x


def preprocess_images_masks(frames_u8, masks, aug=None, out_hw=(48, 64)):
    """One clip. frames (T, H, W, 3) uint8, masks (T, H, W) float32 -> (T, 3, H, W), (T, 1, H, W) float32.
    aug = None or (i, j, th, tw, [brightness, contrast, saturation, hue], order[4])."""
    x = to_tensor(frames_u8)
    m = torch.from_numpy(np.ascontiguousarray(masks, dtype=np.float32)).unsqueeze(1)
    if tuple(x.shape[-2:]) != tuple(out_hw):
        # stored at another size: the dataset's tf.Resize((h, w)) right after ToTensor (:58). Plain bilinear, no
        # antialiasing (torchvision 0.8 / 0.9, see make_golden_dataset.py) = the crop-free case of crop_resize
        hs, ws = x.shape[-2:]
        x = crop_resize(x, 0, 0, hs, ws, *out_hw)
        m = crop_resize(m, 0, 0, hs, ws, *out_hw)
    T, _, H, W = x.shape
    if aug is None:
        return x, m.bool().float()
    i, j, th, tw, factors, order = aug
    x = crop_resize(x, i, j, th, tw, H, W)
    m = crop_resize(m, i, j, th, tw, H, W).bool().float()  # "cast back to 0 or 1 value" (:288-290)
    x = torch.stack([color_jitter(x[t], factors, order) for t in range(T)])
    return x, m


def process_batch(frames_u8, masks, augs=None):
    """Collate (batch-first) + process_batch (time-first): (B, T, H, W, 3) -> (T, B, 3, H, W), (T, B, 1, H, W)."""
    outs = [preprocess_images_masks(frames_u8[b], masks[b], None if augs is None else augs[b])
            for b in range(frames_u8.shape[0])]
    return (torch.stack([o[0] for o in outs]).transpose(1, 0).contiguous(),
            torch.stack([o[1] for o in outs]).transpose(1, 0).contiguous())


def params_to_aug(row):
    """A row of the golden `params` array -> the aug tuple above."""
    return (int(row[0]), int(row[1]), int(row[2]), int(row[3]), [float(v) for v in row[4:8]], [int(v) for v in row[8:12]])
