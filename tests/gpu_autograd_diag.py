"""Diagnostic (not a test): per-parameter gradient difference between the autograd step API, the fused path (time
steps batched) and the fused path run step by step (RAC_TRAIN_PER_STEP=1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_train import _setup  # noqa: E402
from tests.test_gpu_train_autograd import reference_train_step_body  # noqa: E402


def fused(tag, per_step):
    os.environ["RAC_TRAIN_PER_STEP"] = "1" if per_step else "0"
    cfg, sd, model, trainer, batch, ep, eq = _setup(tag, 3)
    trainer.set_noise(ep, eq)
    trainer.forward_backward(batch)
    return {k: trainer.grad_of(k).clone() for k, _ in model.named_parameters()}, (cfg, sd, batch, ep, eq)


def main(tag="vanilla"):
    from robot_aware_control_b200 import SVGConvModel

    ga, (cfg, sd, batch, ep, eq) = fused(tag, False)
    gb, _ = fused(tag, True)
    model = SVGConvModel(cfg).to("cuda")
    model.load_state_dict(sd)
    model.train()
    cfg.batch_size = 4
    opt = torch.optim.Adam(model.parameters(), lr=cfg.lr, betas=(cfg.beta1, 0.999))
    data = {k: v.cuda().float() for k, v in batch.items() if torch.is_tensor(v)}
    reference_train_step_body(cfg, model, opt, data, (ep.cuda(), eq.cuda()), [True] * 3,
                              float(getattr(cfg, "robot_pixel_weight", 0.0)))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    print(f"{'parameter':48s} autograd-vs-batched  autograd-vs-perstep  perstep-vs-batched")
    for k, p in model.named_parameters():
        print(f"{k:48s} {rel(p.grad, ga[k]):.2e}  {rel(p.grad, gb[k]):.2e}  {rel(gb[k], ga[k]):.2e}")


if __name__ == "__main__":
    main(*sys.argv[1:])
