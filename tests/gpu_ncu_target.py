"""Diagnostic (not a test): two SVGConvModel.forward steps at the BASELINE size (g_dim 512, 2000 candidates) -- a short
target for `ncu --set full` captures of single kernels (26 conv_tc launches per step; the prior ConvLSTM pair is
launch 10 / 11 of a step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svg_oracle as so  # noqa: E402
from robot_aware_control_b200 import SVGConvModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfg = so.make_cfg(g_dim=512, z_dim=64)
m = SVGConvModel(cfg)
m.load_state_dict(so.make_state_dict(cfg, 0))
m.eval()
g = torch.Generator().manual_seed(0)
img = torch.rand(n, 3, 48, 64, generator=g).cuda()
act = ((torch.rand(n, 5, generator=g) - 0.5) * 0.1).cuda()
m.init_hidden(n)
for _ in range(2):
    m.forward(img, None, None, None, act)
torch.cuda.synchronize()
print("ok")
