"""Diagnostic (not a test): live CUDA-event time of every convolution layer inside one CEM rollout at the BASELINE
size, via rac_profile_begin/end (events on the launch stream around every launch whose name contains the key).

    python tests/gpu_layer_profile.py [candidates] [--gn] [--ra]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svg_oracle as so  # noqa: E402
from robot_aware_control_b200 import DemoGoalState, State, SVGConvModel, TrajectorySampler, _lib  # noqa: E402

# (name substring, M rows per candidate, N, K) -> algorithmic FLOPs = 2 M N K
LAYERS = [
    ("encoder.c1.0", 3072, 64, 27), ("maxpool.1", 0, 0, 0), ("maxpool.2", 0, 0, 0), ("maxpool.3", 0, 0, 0),
    ("norm_lstm_cell", 0, 0, 0), ("cost_finish", 0, 0, 0),
    ("encoder.c1.1", 3072, 64, 576), ("encoder.c2.0", 768, 128, 576), ("encoder.c2.1", 768, 128, 1152),
    ("encoder.c3.0", 192, 256, 1152), ("encoder.c3.1", 192, 256, 2304), ("encoder.c3.2", 192, 256, 2304),
    ("encoder.c4.0", 48, 512, 2304), ("encoder.c4.1", 48, 512, 4608), ("encoder.c4.2", 48, 512, 4608),
    ("prior_input_conv", 48, 512, 4653), ("prior.lstm.0", 48, 2048, 25600), ("prior.lstm.1", 48, 2048, 9216),
    ("prior.mu_net", 48, 128, 4608), ("frame_pred_input_conv", 48, 512, 5229),
    ("frame_predictor.lstm.0", 48, 2048, 25600), ("frame_predictor.lstm.1", 48, 2048, 9216),
    ("decoder.upc2.0", 48, 512, 4608), ("decoder.upc2.1", 48, 512, 4608), ("decoder.upc2.2", 48, 256, 4608),
    ("decoder.upc3.0", 192, 256, 4608), ("decoder.upc3.1", 192, 256, 2304), ("decoder.upc3.2", 192, 128, 2304),
    ("decoder.upc4.0", 768, 128, 2304), ("decoder.upc4.1", 768, 64, 1152), ("decoder.upc5.0", 3072, 64, 1152),
    ("decoder.upc5.1", 3072, 4, 576),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2000
    gn = "--gn" in sys.argv
    L = 5
    ra = "--ra" in sys.argv  # robot-aware model: mask + future mask + robot state, dontcare cost (BASELINE configs[4])
    kw = dict(model_use_mask=True, model_use_future_mask=True, model_use_robot_state=True,
              reconstruction_loss="dontcare_l1", reward_type="dontcare") if ra else {}
    cfg = so.make_cfg(g_dim=512, z_dim=64, lstm_group_norm=gn, **kw)
    model = SVGConvModel(cfg)
    model.load_state_dict(so.make_state_dict(cfg, 0))
    model.eval()
    rs = np.random.RandomState(0)
    start = State(img=rs.randint(0, 256, (48, 64, 3)).astype(np.uint8))
    goal = DemoGoalState(imgs=[rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)], masks=[np.zeros((1, 48, 64), np.float32)])
    ts = TrajectorySampler(cfg, model)
    g = torch.Generator().manual_seed(0)
    actions = torch.cat([(torch.rand(n, L, 2, generator=g) - 0.5) * 0.1, torch.zeros(n, L, 3)], 2).cuda()
    lib = _lib.load()
    states = masks = None
    if ra:
        states = torch.rand(L + 1, n, 5, generator=g).cuda()
        masks = torch.zeros(L + 1, n, 1, 48, 64)
        masks[:, :, :, 10:30, 20:44] = 1
        masks = masks.cuda()
    rollout = lambda: ts.generate_model_rollouts(actions, start, goal, states=states, masks=masks)
    for _ in range(2):
        rollout()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rollout()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    print(f"rollout {n} x {L}: {total:.2f} ms = {total / L:.3f} ms/step, {n * L / total * 1e3:.0f} frames/s")
    acc = 0.0
    for name, M, N, K in LAYERS:
        if name == "norm_lstm_cell" and not gn:
            continue
        _lib.check(lib.rac_profile_begin(model.handle, name.encode(), 64), model.handle, "begin")
        rollout()
        pl, pms = C.c_int64(), C.c_double()
        _lib.check(lib.rac_profile_end(model.handle, C.byref(pl), C.byref(pms)), model.handle, "end")
        per_step = pms.value / L
        acc += per_step
        fl = 2.0 * M * N * K * n
        print(f"{name:28s} launches/step {pl.value / L:4.1f}  {per_step:7.3f} ms/step  {100 * per_step * L / total:5.1f}%  "
              f"{fl / (per_step * 1e-3) / 1e12:7.0f} TFLOP/s algorithmic")
    print(f"sum of timed convs {acc:.3f} ms/step of {total / L:.3f}")


if __name__ == "__main__":
    main()
