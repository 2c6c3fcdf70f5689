"""TEST INFRASTRUCTURE -- golden vectors at the BASELINE.json model size (g_dim 512, z_dim 64, action_dim 5): runs the
UNMODIFIED reference (/root/reference through oracle/ref_shim.py) on CPU and stores its outputs under tests/golden.

    python -m oracle.make_golden_g512 [forward] [rollout] [trained] [train]

* forward_g512_{vanilla,ra}.npz  SVGConvModel.forward, 2 recurrent steps (+ future mask / future state for `ra`)
* rollout_g512_{vanilla,ra,ra_sparse}.npz  generate_model_rollouts, 5 noisy steps (eps supplied), frames + fp64 costs
* rollout_g512_trained_{vanilla,ra}.npz  the same with a "trained-like" weight set (svg_oracle.trained_like: LSTM gate
  pre-activations of magnitude 3-5 and a decoder head with real contrast, instead of the forgiving random init)
* train_g512_{l1,config3}.npz  two consecutive PredictionTrainer._train_step calls at batch 16 / n_future 5:
  BASELINE configs[0] (l1, vanilla) and the configs[3] shape (dontcare_l1, mask + future mask + robot state,
  scheduled sampling with the model's own frame at i > 1)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, svg_oracle as so  # noqa: E402
from oracle.make_golden import EpsFeeder, FakeRobotModel, inputs_forward, synth_masks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
G_DIM, Z_DIM = 512, 64
ROLL_N, ROLL_L = 3, 5
TRAIN_B, TRAIN_T = 16, 6


def rollout_inputs(seed, N, L, z_dim):
    g = torch.Generator().manual_seed(seed)
    actions = torch.cat([(torch.rand(N, L, 2, generator=g) - 0.5) * 0.1, torch.zeros(N, L, 3)], 2)
    eps = torch.randn(L, N, z_dim, 6, 8, generator=g)
    states = torch.rand(L + 1, N, 5, generator=g)
    return actions, eps, states


def train_batch(seed, cfg, robot_aware, B=TRAIN_B, T=TRAIN_T):
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
    only = set(sys.argv[1:]) or {"forward", "rollout", "trained", "train"}
    os.makedirs(OUT, exist_ok=True)
    mods = ref_shim.import_reference()
    dyn = mods["src.prediction.models.dynamics"]
    lstm_mod = sys.modules["src.prediction.models.lstm"]
    ts_mod = mods["src.cem.trajectory_sampler"]
    losses = mods["src.prediction.losses"]
    State, DemoGoalState = mods["src.utils.state"].State, mods["src.utils.state"].DemoGoalState
    feeder = EpsFeeder()
    lstm_mod.GaussianConvLSTM.reparameterize = lambda self, mu, logvar: feeder(self, mu, logvar)

    def build(cfg, sd):
        torch.manual_seed(1234)
        m = dyn.SVGConvModel(cfg)
        res = m.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        m.eval()
        return m

    # ------------------------------------------------------------------ forward, 2 steps
    if "forward" in only:
        for tag, kw in (("vanilla", dict(robot_aware=False)),
                        ("ra", dict(robot_aware=True, future_mask=True, future_robot_state=True))):
            cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, **kw)
            model = build(cfg, so.make_state_dict(cfg, 111))
            B = 2
            d = inputs_forward(121, B, cfg)
            model.init_hidden(B)
            out = {}
            with torch.no_grad():
                for t in range(2):
                    mask = robot = None
                    if cfg.model_use_mask:
                        mask = torch.cat([d["mask"][t], d["mask"][t + 1]], 1)
                    if cfg.model_use_robot_state:
                        robot = (d["robot"][t], d["robot"][t + 1])
                    feeder.queue = [d["eps"][t].clone()]
                    x_pred, skip, _, _, mu_p, logvar_p = model.forward(d["image"][t], mask, robot, None, d["action"][t])
                    out[f"x_pred{t}"] = x_pred.numpy()
                    out[f"mu_p{t}"] = mu_p.numpy()
                    out[f"logvar_p{t}"] = logvar_p.numpy()
            np.savez_compressed(os.path.join(OUT, f"forward_g512_{tag}.npz"), weight_seed=111, input_seed=121, B=B, **out)
            print("forward g512", tag, float(out["x_pred1"].std()))

    # ------------------------------------------------------------------ rollouts, 5 noisy steps
    scene = np.load(os.path.join(OUT, "scene.npz"))
    start_img, goal_imgs, goal_masks = scene["start_img"], list(scene["goal_imgs"]), list(scene["goal_masks"])
    N, L = ROLL_N, ROLL_L
    jobs = []
    if "rollout" in only:
        jobs += [("vanilla", dict(robot_aware=False), (), False), ("ra", dict(robot_aware=True, future_mask=True), (), False),
                 ("ra_sparse", dict(robot_aware=True), ("--sparse_cost", "True"), False)]
    if "trained" in only:
        jobs += [("trained_vanilla", dict(robot_aware=False), (), True),
                 ("trained_ra", dict(robot_aware=True, future_mask=True), (), True)]
    for tag, kw, extra, trained in jobs:
        cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM,
                                extra=("--candidates_batch_size", str(N), "--topk", str(N)) + tuple(extra), **kw)
        sd = so.make_state_dict(cfg, 112)
        if trained:
            sd = so.trained_like(sd)
        model = build(cfg, sd)
        actions, eps, states = rollout_inputs(131, N, L, cfg.z_dim)
        masks = synth_masks(141, L, N)
        sampler = ts_mod.TrajectorySampler.__new__(ts_mod.TrajectorySampler)
        sampler.cfg, sampler.model, sampler.cost = cfg, model, losses.RobotWorldCost(cfg)
        sampler.low = torch.from_numpy(np.array([[0.015, -0.3, 0.1, 0, 0]], dtype=np.float32))
        sampler.high = torch.from_numpy(np.array([[0.55, 0.3, 0.4, 1, 1]], dtype=np.float32))
        sampler.robot_model = FakeRobotModel(states, masks)
        feeder.queue = [eps[t].clone() for t in range(L)]
        start = State(img=start_img, state=np.array([0.3, 0.0, 0.2, 0.0, 0.0], dtype=np.float32), qpos=np.zeros(6))
        goal = DemoGoalState(imgs=goal_imgs, masks=goal_masks)
        r = sampler.generate_model_rollouts(actions, start, goal, ret_obs=True)
        inv = np.empty(N, dtype=np.int64)
        inv[r["topk_idx"]] = np.arange(N)
        obs = r["obs"][inv]
        np.savez_compressed(os.path.join(OUT, f"rollout_g512_{tag}.npz"), weight_seed=112, input_seed=131, mask_seed=141,
                            N=N, L=L, trained=int(trained), sum_cost=r["sum_cost"], obs=obs.astype(np.float32))
        print("rollout g512", tag, r["sum_cost"], "frame std", float(obs[:, -1].std()))

    # ------------------------------------------------------------------ training, batch 16 / n_future 5
    if "train" in only:
        import importlib

        trainer_mod = importlib.import_module("src.prediction.trainer")
        B, T = TRAIN_B, TRAIN_T
        for tag, kw in (("l1", dict(robot_aware=False)), ("config3", dict(robot_aware=True, future_mask=True))):
            cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, extra=("--lr", "1e-4", "--beta", "1e-4"), **kw)
            cfg.multiview = False
            sd = so.make_state_dict(cfg, 117)
            tr = trainer_mod.PredictionTrainer.__new__(trainer_mod.PredictionTrainer)
            tr._config, tr._device = cfg, cfg.device
            torch.manual_seed(0)
            tr._init_models(cfg)
            tr.model.load_state_dict(sd)
            tr._scheduled_sampling = False
            if tag == "config3":  # scheduled sampling: the model's own frame at every step i > 1 (worst case for parity)
                tr._scheduled_sampling = True
                tr._use_true_token = lambda: False
            tr._step = 0
            tr.model.train()
            batch, eps_p, eps_q = train_batch(123, cfg, kw["robot_aware"])
            batch_ref = dict(batch, qpos=torch.zeros(T, B, 6), robot=["sawyer"] * B, folder=["x"] * B)
            out = {}
            for step in range(2):
                feeder.queue = []
                for t in range(T - 1):
                    feeder.queue += [eps_p[t].clone(), eps_q[t].clone()]
                ls = tr._train_step(batch_ref)
                grads = {k: p.grad.detach().clone() for k, p in tr.model.named_parameters()}
                keys, gn, gs = summarize(grads)
                params = {k: p.detach().clone() for k, p in tr.model.named_parameters()}
                _, pn, ps = summarize(params)
                bufs = {k: v.detach().clone().float() for k, v in tr.model.named_buffers() if "running" in k}
                bkeys, bn, _ = summarize(bufs)
                out[f"recon{step}"] = ls["recon_loss"] * cfg.n_future
                out[f"kld{step}"] = ls["kld"] * cfg.n_future
                out[f"robot{step}"] = ls["robot_loss"] * cfg.n_future
                out[f"world{step}"] = ls["world_loss"] * cfg.n_future
                out[f"grad_norm{step}"], out[f"grad_sample{step}"] = gn, gs
                out[f"param_norm{step}"], out[f"param_sample{step}"] = pn, ps
                out[f"running_norm{step}"] = bn
                print("train g512", tag, step, out[f"recon{step}"], out[f"kld{step}"])
            np.savez_compressed(os.path.join(OUT, f"train_g512_{tag}.npz"), weight_seed=117, input_seed=123, B=B, T=T,
                                lr=1e-4, beta=1e-4, keys=np.array(keys), running_keys=np.array(bkeys), **out)


if __name__ == "__main__":
    main()
