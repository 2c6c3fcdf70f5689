"""Evaluation metrics with the reference's signatures (src/utils/metrics.py:45-78 `ssim`, `psnr`;
src/prediction/losses.py:80-94 `world_psnr_criterion`), computed by libracb200.so (csrc/metric_kernels.cu).
`mask=` is an extension: robot pixels of both images are zeroed inside the kernel (what `_eval_step` does with
zero_robot_region before calling the metrics, trainer.py:685-690), so the blacked copies are never materialised."""
import torch

from . import _lib


def _dev(t):
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


@torch.no_grad()
def psnr(estimates, targets, data_dims=3, mask=None, clamp01=False):
    """metrics.py:57-78. Returns a (B,) CUDA tensor. The reference drops into a debugger when a value leaves
    [-0.01, 1.01]; that check is not reproduced."""
    if data_dims != 3 or estimates.dim() != 4:
        raise NotImplementedError("psnr: (B, C, H, W) inputs with data_dims=3 (the only use in the reference)")
    e, t = _dev(estimates), _dev(targets)
    m = _dev(mask) if mask is not None else None
    n, c, h, w = e.shape
    out = torch.empty(n, device="cuda")
    if n == 0:
        return out
    _lib.check(_lib.load().rac_psnr(_lib.ptr(e), _lib.ptr(t), _lib.ptr(m), int(bool(clamp01)), _lib.ptr(out), n, c,
                                    h * w, _lib.stream_ptr()), None, "rac_psnr")
    return out


@torch.no_grad()
def world_psnr_criterion(prediction, target, mask):
    """losses.py:80-94. Returns a (B,) CUDA tensor."""
    p, t, m = _dev(prediction), _dev(target), _dev(mask)
    n, c, h, w = p.shape
    assert c == 3
    out = torch.empty(n, device="cuda")
    _lib.check(_lib.load().rac_world_psnr(_lib.ptr(p), _lib.ptr(t), _lib.ptr(m), _lib.ptr(out), n, h * w,
                                          _lib.stream_ptr()), None, "rac_world_psnr")
    return out


def _ssim(img1, img2, mask, want_map):
    a, b = _dev(img1), _dev(img2)
    m = _dev(mask) if mask is not None else None
    n, c, h, w = a.shape
    smap = torch.empty(n, c, h, w, device="cuda") if want_map else None
    means = torch.empty(n * c, device="cuda")
    _lib.check(_lib.load().rac_ssim(_lib.ptr(a), _lib.ptr(b), _lib.ptr(m), _lib.ptr(smap), _lib.ptr(means), n, c, h, w,
                                    _lib.stream_ptr()), None, "rac_ssim")
    return smap, means


@torch.no_grad()
def ssim(img1, img2, window_size=11, mask=None):
    """metrics.py:45-54: the SSIM map as a numpy array (B, C, H, W)."""
    if window_size != 11:
        raise NotImplementedError("ssim: window_size 11 (the reference default, the only one it uses)")
    return _ssim(img1, img2, mask, True)[0].cpu().numpy()


@torch.no_grad()
def ssim_mean(img1, img2, mask=None):
    """ssim(...).mean() without the map round trip: CUDA scalar."""
    return _ssim(img1, img2, mask, False)[1].mean()
