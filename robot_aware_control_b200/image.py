"""zero_robot_region (reference src/utils/image.py:5-20). Host-side convenience with the reference signature; inside
the rollout the same operation is fused into the frame epilogue (csrc/epilogue.cuh::epi_frame)."""
import numpy as np
import torch


def zero_robot_region(mask, image, inplace=False):
    if isinstance(mask, torch.Tensor):
        keep = (~mask.type(torch.bool)).to(image.dtype)  # (B,1,H,W) broadcasts over the 3 colour channels
        if inplace:
            return image.mul_(keep)
        return image * keep
    robot_mask = np.asarray(mask).astype(bool)
    if not inplace:
        image = image.copy()
    image[robot_mask] = 0
    return image
