"""TEST INFRASTRUCTURE -- per-parameter comparison of the CUDA training step with the CPU training oracle on a B200:
python -m tests.gpu_train_diag [vanilla|ra|ra_gn] [simt|tc]"""
import sys

import numpy as np
import torch

from oracle import svg_oracle as so
from oracle.make_golden import G_DIM, Z_DIM
from oracle.make_golden_train import make_batch
from oracle.train_oracle import TrainOracle


def setup(tag, impl="tc", lr=1e-3, beta=1e-2):
    from robot_aware_control_b200 import SVGConvModel, SVGTrainer

    if tag == "vanilla":
        cfg = so.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, lr=lr, beta=beta, beta1=0.9, n_future=3, n_past=1)
    else:
        cfg = so.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, model_use_mask=True, model_use_future_mask=True,
                          model_use_robot_state=True, reconstruction_loss="dontcare_l1", reward_type="dontcare",
                          lstm_group_norm=tag.endswith("_gn"), lr=lr, beta=beta, beta1=0.9, n_future=3, n_past=1)
    sd = so.make_state_dict(cfg, 17)
    model = SVGConvModel(cfg, conv_impl=1 if impl == "simt" else 0)
    model.load_state_dict(sd)
    model.train()
    trainer = SVGTrainer(cfg, model)
    oracle = TrainOracle(cfg, sd, lr=lr, beta=beta)
    batch, eps_p, eps_q = make_batch(23, cfg, tag != "vanilla")
    return cfg, model, trainer, oracle, batch, eps_p, eps_q


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "vanilla"
    impl = sys.argv[2] if len(sys.argv) > 2 else "tc"
    cfg, model, trainer, oracle, batch, eps_p, eps_q = setup(tag, impl)
    oracle.model.emulate_bf16 = "fp32" not in sys.argv  # default: the oracle rounds where the CUDA path stores bf16
    info, ref = oracle.loss_and_grads(batch, eps_p, eps_q)
    trainer.set_noise(eps_p, eps_q)
    losses = trainer.forward_backward(batch).cpu().numpy()
    torch.cuda.synchronize()
    print(f"recon gpu {losses[0]:.6f} ref {info['recon_loss']:.6f} | kld gpu {losses[1]:.6f} ref {info['kld']:.6f}")
    worst = []
    for k in oracle.param_keys:
        g = trainer.grad_of(k).cpu()
        r = ref[k]
        rel = float((g - r).norm() / (r.norm() + 1e-12))
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-20))
        worst.append((rel, k, float(r.norm()), float(g.norm()), cos))
    for rel, k, rn, gn, cos in sorted(worst, reverse=True)[::(1 if tag.endswith("_gn") else 3)]:
        print(f"{k:45s} rel_err {rel:9.4f}  |ref| {rn:10.4e} |gpu| {gn:10.4e} cos {cos:7.4f}")
    rels = np.array([w[0] for w in worst])
    print("median rel err", np.median(rels), "max", rels.max(), "n", len(rels))


if __name__ == "__main__":
    main()
