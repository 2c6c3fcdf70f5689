"""Pins the CPU oracle at the BASELINE.json model size (g_dim 512, z_dim 64 -- every z channel live) against outputs of
the UNMODIFIED reference (tests/golden/*_g512_*.npz from oracle/make_golden_g512.py): forward, 5-step noisy rollouts
(vanilla / robot-aware / sparse, random-init and "trained-like" weights) and one batch-16 training step. CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import svg_oracle as so
from oracle.make_golden import inputs_forward, synth_masks
from oracle.make_golden_g512 import G_DIM, Z_DIM, rollout_inputs, summarize, train_batch
from oracle.train_oracle import TrainOracle


def cfg_for(tag, **kw):
    if "vanilla" in tag or tag == "l1":
        return so.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, **kw)
    return so.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, model_use_mask=True, model_use_robot_state=True,
                       reconstruction_loss="dontcare_l1", reward_type="dontcare", **kw)


def rollout_setup(golden_dir, tag):
    gold = np.load(os.path.join(golden_dir, f"rollout_g512_{tag}.npz"))
    extra = {}
    if tag in ("ra", "trained_ra"):
        extra = dict(model_use_future_mask=True)
    if tag == "ra_sparse":
        extra = dict(sparse_cost=True)
    N, L = int(gold["N"]), int(gold["L"])
    cfg = cfg_for(tag, topk=N, **extra)
    sd = so.make_state_dict(cfg, int(gold["weight_seed"]))
    if int(gold["trained"]):
        sd = so.trained_like(sd)
    actions, eps, states = rollout_inputs(int(gold["input_seed"]), N, L, cfg.z_dim)
    masks = synth_masks(int(gold["mask_seed"]), L, N)
    return gold, cfg, sd, actions, eps, states, masks


@pytest.mark.parametrize("tag", ["vanilla", "ra"])
def test_forward_g512_matches_reference(golden_dir, tag):
    gold = np.load(os.path.join(golden_dir, f"forward_g512_{tag}.npz"))
    extra = dict(model_use_future_mask=True, model_use_future_robot_state=True) if tag == "ra" else {}
    cfg = cfg_for(tag, **extra)
    model = so.SVGOracle(cfg, so.make_state_dict(cfg, int(gold["weight_seed"])))
    B = int(gold["B"])
    d = inputs_forward(int(gold["input_seed"]), B, cfg)
    model.init_hidden(B)
    for t in range(2):
        mask = torch.cat([d["mask"][t], d["mask"][t + 1]], 1) if cfg.model_use_mask else None
        robot = (d["robot"][t], d["robot"][t + 1]) if cfg.model_use_robot_state else None
        x_pred, _, _, _, mu_p, logvar_p = model.forward(d["image"][t], mask, robot, d["action"][t], d["eps"][t])
        np.testing.assert_allclose(x_pred.numpy(), gold[f"x_pred{t}"], rtol=1e-5, atol=3e-6)
        np.testing.assert_allclose(mu_p.numpy(), gold[f"mu_p{t}"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(logvar_p.numpy(), gold[f"logvar_p{t}"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("tag", ["vanilla", "ra", "ra_sparse", "trained_vanilla", "trained_ra"])
def test_rollout_g512_matches_reference(golden_dir, tag):
    gold, cfg, sd, actions, eps, states, masks = rollout_setup(golden_dir, tag)
    scene = np.load(os.path.join(golden_dir, "scene.npz"))
    model = so.SVGOracle(cfg, sd)
    r = so.rollout_cost(model, cfg, actions, scene["start_img"], list(scene["goal_imgs"]), list(scene["goal_masks"]),
                        states, masks, eps, ret_obs=True)
    np.testing.assert_allclose(r["obs"], gold["obs"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(r["sum_cost"], gold["sum_cost"], rtol=5e-6)


def test_train_step_g512_matches_reference(golden_dir):
    """BASELINE configs[0]: batch 16, n_past 1 / n_future 5, l1 -- one step of the oracle (the second reference step
    and the configs[3] shape are checked on the GPU against the same fixtures)."""
    gold = np.load(os.path.join(golden_dir, "train_g512_l1.npz"))
    cfg = cfg_for("l1")
    tr = TrainOracle(cfg, so.make_state_dict(cfg, int(gold["weight_seed"])), lr=float(gold["lr"]), beta=float(gold["beta"]))
    batch, eps_p, eps_q = train_batch(int(gold["input_seed"]), cfg, False)
    info, grads = tr.train_step(batch, eps_p, eps_q)
    np.testing.assert_allclose(info["recon_loss"], gold["recon0"], rtol=2e-5)
    np.testing.assert_allclose(info["kld"], gold["kld0"], rtol=2e-4)
    keys, gn, _ = summarize(grads)
    assert keys == list(gold["keys"])
    np.testing.assert_allclose(gn, gold["grad_norm0"], rtol=3e-3, atol=1e-7)
    _, pn, _ = summarize({k: tr.model.sd[k].detach() for k in tr.param_keys})
    np.testing.assert_allclose(pn, gold["param_norm0"], rtol=1e-5)
