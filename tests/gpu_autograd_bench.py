"""Timing (not a test): the reference-style training step -- tests/test_gpu_train_autograd.py::reference_train_step_body,
i.e. train-mode SVGConvModel.forward under torch autograd + torch losses + torch.optim.Adam, with the reference's
per-frame `.cpu().item()` logging syncs -- next to the fused SVGTrainer.train_step, at BASELINE configs[0] shape
(batch 16, n_past 1 / n_future 5, g512 / z64). Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svg_oracle as so  # noqa: E402
from robot_aware_control_b200 import SVGConvModel, SVGTrainer  # noqa: E402
from tests.test_gpu_train_autograd import reference_train_step_body  # noqa: E402


def main(steps=5, warmup=3):
    dev = torch.device("cuda")
    cfg = so.make_cfg(g_dim=512, z_dim=64, action_dim=5, lr=1e-4, beta=1e-4, beta1=0.9, n_future=5, n_past=1)
    cfg.batch_size = 16
    sd = so.make_state_dict(cfg, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    T, B = 6, 16
    data = {"images": torch.rand(T, B, 3, 48, 64, device=dev, generator=g),
            "actions": torch.rand(T - 1, B, 5, device=dev, generator=g) * 0.1 - 0.05,
            "states": torch.rand(T, B, 5, device=dev, generator=g), "masks": torch.zeros(T, B, 1, 48, 64, device=dev)}
    noise = (torch.randn(T - 1, B, 64, 6, 8, device=dev, generator=g), torch.randn(T - 1, B, 64, 6, 8, device=dev, generator=g))
    out = {}
    for mode in ("autograd", "fused"):
        model = SVGConvModel(cfg).to(dev)
        model.load_state_dict(sd)
        model.train()
        if mode == "autograd":
            opt = torch.optim.Adam(model.parameters(), lr=cfg.lr, betas=(cfg.beta1, 0.999))
            step = lambda: reference_train_step_body(cfg, model, opt, data, noise, [True] * 5, 0.0)
        else:
            trainer = SVGTrainer(cfg, model)
            step = lambda: trainer.train_step(data)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[mode + "_ms_per_step"] = e0.elapsed_time(e1) / steps
        del model
        torch.cuda.empty_cache()
    out["workload"] = "SVG training step, batch 16, n_past 1 / n_future 5, g_dim 512 z_dim 64, l1, Adam; reference-style loop vs fused C step"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
