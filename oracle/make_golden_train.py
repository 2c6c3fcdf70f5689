"""TEST INFRASTRUCTURE -- golden vectors of the training step: runs the UNMODIFIED reference
PredictionTrainer._train_step (src/prediction/trainer.py:326-465) on CPU with deterministic synthetic weights and
injected reparameterisation noise, and stores losses, per-parameter gradient norms / samples and post-Adam parameter
samples in tests/golden/train_*.npz.   python -m oracle.make_golden_train"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, svg_oracle as so  # noqa: E402
from oracle.make_golden import EpsFeeder, G_DIM, Z_DIM, synth_masks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
B, T = 4, 4  # n_past 1 + n_future 3
HIGH_MOVEMENT = torch.tensor([True, False, True, False])  # data["high_movement"] of the "ra_bw" case (trainer.py:426-429)


def make_batch(seed, cfg, robot_aware):
    g = torch.Generator().manual_seed(seed)
    batch = {
        "images": torch.rand(T, B, 3, 48, 64, generator=g),
        "actions": (torch.rand(T - 1, B, cfg.action_dim, generator=g) - 0.5) * 0.1,
        "states": torch.rand(T, B, 5, generator=g),
        "masks": synth_masks(seed + 1, T - 1, B) if robot_aware else torch.zeros(T, B, 1, 48, 64),
    }
    eps_prior = torch.randn(T - 1, B, cfg.z_dim, 6, 8, generator=g)
    eps_post = torch.randn(T - 1, B, cfg.z_dim, 6, 8, generator=g)
    return batch, eps_prior, eps_post


def summarize(named):
    keys = sorted(named)
    norms = np.array([float(named[k].double().norm()) for k in keys])
    samples = np.concatenate([named[k].reshape(-1)[:4].double().numpy() for k in keys])
    return keys, norms, samples


def main():
    mods = ref_shim.import_reference()
    lstm_mod = sys.modules["src.prediction.models.lstm"]
    import importlib

    trainer_mod = importlib.import_module("src.prediction.trainer")
    feeder = EpsFeeder()
    lstm_mod.GaussianConvLSTM.reparameterize = lambda self, mu, logvar: feeder(self, mu, logvar)
    only = sys.argv[1:]  # optional: tags to (re)generate; default all
    for tag, kw in (("vanilla", dict(robot_aware=False)), ("ra", dict(robot_aware=True, future_mask=True)),
                    ("ra_sampled", dict(robot_aware=True, future_mask=True)),
                    ("ra_fixedskip", dict(robot_aware=True, future_mask=True)),
                    ("vanilla_fixedskip_sampled", dict(robot_aware=False)),
                    ("ra_gn", dict(robot_aware=True, future_mask=True)),
                    # the remaining cfg.reconstruction_loss kinds (trainer.py:149-161) and the movement weighting
                    ("vanilla_mse", dict(robot_aware=False)), ("ra_dcmse", dict(robot_aware=True, future_mask=True)),
                    ("ra_bw", dict(robot_aware=True, future_mask=True))):
        if only and tag not in only:
            continue
        extra = ("--n_future", str(T - 1), "--batch_size", str(B), "--lr", "1e-3", "--beta", "1e-2")
        if tag == "vanilla_mse":
            extra += ("--reconstruction_loss", "mse")
        if tag == "ra_dcmse":
            extra += ("--reconstruction_loss", "dontcare_mse", "--robot_pixel_weight", "0.25")
        if tag == "ra_bw":
            extra += ("--load_movement_info", "True", "--movement_weight", "3.0")
        if "fixedskip" in tag:  # the config default (src/config/__init__.py:217-222): decoder skips of the first frame
            extra += ("--last_frame_skip", "False")
        if tag.endswith("_gn"):  # NormConvLSTMCell (lstm.py:151-198), the cell of the authors' deployed checkpoints
            extra += ("--lstm_group_norm", "True")
        cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, extra=extra, **kw)
        cfg.multiview = False
        sd = so.make_state_dict(cfg, 17)
        tr = trainer_mod.PredictionTrainer.__new__(trainer_mod.PredictionTrainer)
        tr._config, tr._device = cfg, cfg.device
        torch.manual_seed(0)
        tr._init_models(cfg)
        tr.model.load_state_dict(sd)
        tr._scheduled_sampling = False
        if tag.endswith("sampled"):  # scheduled sampling with the model's own frame at every step i > 1
            tr._scheduled_sampling = True
            tr._use_true_token = lambda: False
        tr._step = 0
        tr.model.train()
        kw = dict(kw)
        batch, eps_p, eps_q = make_batch(23, cfg, kw["robot_aware"])
        batch_ref = dict(batch, qpos=torch.zeros(T, B, 6), robot=["sawyer"] * B, folder=["x"] * B)
        if tag == "ra_bw":
            batch_ref["high_movement"] = HIGH_MOVEMENT.clone()
        out = {}
        for step in range(2):
            feeder.queue = []
            for t in range(T - 1):
                feeder.queue += [eps_p[t].clone(), eps_q[t].clone()]
            losses = tr._train_step(batch_ref)
            grads = {k: p.grad.detach().clone() for k, p in tr.model.named_parameters()}
            keys, gn, gs = summarize(grads)
            params = {k: p.detach().clone() for k, p in tr.model.named_parameters()}
            _, pn, ps = summarize(params)
            bufs = {k: v.detach().clone().float() for k, v in tr.model.named_buffers() if "running" in k}
            bkeys, bn, _ = summarize(bufs)
            out[f"recon{step}"] = losses["recon_loss"] * cfg.n_future  # undo the logging average (trainer.py:463-464)
            out[f"kld{step}"] = losses["kld"] * cfg.n_future
            out[f"robot{step}"] = losses["robot_loss"] * cfg.n_future  # logged metrics (trainer.py:436-439)
            out[f"world{step}"] = losses["world_loss"] * cfg.n_future
            out[f"grad_norm{step}"], out[f"grad_sample{step}"] = gn, gs
            out[f"param_norm{step}"], out[f"param_sample{step}"] = pn, ps
            out[f"running_norm{step}"] = bn
            print(tag, step, out[f"recon{step}"], out[f"kld{step}"])
        np.savez_compressed(os.path.join(OUT, f"train_{tag}.npz"), weight_seed=17, input_seed=23, B=B, T=T, lr=1e-3,
                            beta=1e-2, keys=np.array(keys), running_keys=np.array(bkeys), **out)


if __name__ == "__main__":
    main()
