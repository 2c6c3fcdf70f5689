"""Diagnostic (not a test): launch time of rac_process_batch (the loader's per-clip preprocessing + augmentation + time-first
transposition) against its HBM floor: per pixel 3 B (uint8 rgb) + 4 B (fp32 mask) read, 12 B + 4 B
written = 23 B algorithmic with fp32 masks (29 B was the round-1 figure counting the uint8 mask variant separately).

    python tests/gpu_data_bench.py [clips]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robot_aware_control_b200 import data as D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = 6
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (B, T, 48, 64, 3), dtype=torch.uint8, generator=g).cuda()
masks = (torch.rand(B, T, 48, 64, generator=g) > 0.8).float().cuda()
res = {}
for name, augment in (("to_tensor_only", False), ("augmented", True)):
    aug = None
    if augment:
        import random

        random.seed(0)
        torch.manual_seed(0)
        aug = [D.sample_augment(48, 64) for _ in range(B)]
    run = lambda: D.preprocess_clips(frames, masks, aug)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    pix = B * T * 48 * 64
    res[name] = {"ms_per_batch": ms, "clips": B, "frames": B * T, "algorithmic_bytes": pix * 23,
                 "achieved_gbs": pix * 23 / (ms * 1e-3) / 1e9}
print(json.dumps(res))
