"""Diagnostic (not a test): the stand-alone robot-aware planning cost kernel (rac_masked_cost, ImgDontcareCost in the
reference NCHW fp32 layout) on two alternating 805 MB input sets (> 126 MB L2) -- a target for
`ncu --set full -k regex:masked_cost` (dram__bytes_read.sum vs the 49 156 algorithmic bytes per candidate)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robot_aware_control_b200 import _lib  # noqa: E402

n = 16384
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
sets = [(torch.rand(n, 3, 48, 64, device="cuda", generator=g), (torch.rand(n, 1, 48, 64, device="cuda", generator=g) > 0.8).float())
        for _ in range(2)]
goal = torch.rand(3, 48, 64, device="cuda", generator=g)
gmask = (torch.rand(1, 48, 64, device="cuda", generator=g) > 0.8).float()
out = torch.empty(n, device="cuda")
for i in range(6):
    c, m = sets[i & 1]
    lib.rac_masked_cost(_lib.ptr(c), _lib.ptr(goal), _lib.ptr(m), _lib.ptr(gmask), 1, _lib.ptr(out), n, 48 * 64, _lib.stream_ptr())
torch.cuda.synchronize()
print("ok", float(out.sum()))
