"""GPU parity at the BASELINE.json model size (g_dim 512, z_dim 64, action_dim 5) against outputs of the UNMODIFIED
reference (tests/golden/*_g512_*.npz, oracle/make_golden_g512.py), through the C ABI:

* forward, two recurrent steps, vanilla and robot-aware (+ future mask / future state)
* 5-step autoregressive rollouts WITH z noise -- the bf16-accumulation worst case -- vanilla, robot-aware (configs[4]),
  robot-aware sparse; random-init and "trained-like" weights (saturating LSTM gates, contrasty decoder head)
* the training step at batch 16 / n_future 5: BASELINE configs[0] (l1) and the configs[3] shape (dontcare_l1,
  mask + future mask + robot state, scheduled sampling), two consecutive steps

Tolerances (north_star): frames within 1e-2 max-abs on [0,1] pixels; costs relative (they carry 255 * sqrt(#pixels))."""
import os

import numpy as np
import pytest
import torch

from oracle import svg_oracle as so
from oracle.make_golden import inputs_forward
from oracle.make_golden_g512 import train_batch
from tests.test_oracle_golden_g512 import cfg_for, rollout_setup

pytestmark = pytest.mark.gpu
PIX_TOL = 1e-2


def _model(cfg, sd):
    from robot_aware_control_b200 import SVGConvModel

    m = SVGConvModel(cfg)
    m.load_state_dict(sd)
    m.eval()
    return m


@pytest.mark.parametrize("tag", ["vanilla", "ra"])
def test_forward_g512_matches_reference_golden(golden_dir, tag):
    gold = np.load(os.path.join(golden_dir, f"forward_g512_{tag}.npz"))
    extra = dict(model_use_future_mask=True, model_use_future_robot_state=True) if tag == "ra" else {}
    cfg = cfg_for(tag, **extra)
    m = _model(cfg, so.make_state_dict(cfg, int(gold["weight_seed"])))
    B = int(gold["B"])
    d = inputs_forward(int(gold["input_seed"]), B, cfg)
    m.init_hidden(B)
    for t in range(2):
        mask = torch.cat([d["mask"][t], d["mask"][t + 1]], 1) if cfg.model_use_mask else None
        robot = (d["robot"][t], d["robot"][t + 1]) if cfg.model_use_robot_state else None
        m.set_noise(eps=d["eps"][t])
        x_pred, _, _, _, mu_p, logvar_p = m.forward(d["image"][t], mask, robot, None, d["action"][t])
        assert np.abs(x_pred.cpu().numpy() - gold[f"x_pred{t}"]).max() < PIX_TOL
        assert np.abs(mu_p.cpu().numpy() - gold[f"mu_p{t}"]).max() < 5e-2
        assert np.abs(logvar_p.cpu().numpy() - gold[f"logvar_p{t}"]).max() < 5e-2


@pytest.mark.parametrize("tag", ["vanilla", "ra", "ra_sparse", "trained_vanilla", "trained_ra"])
def test_rollout_g512_noisy_5_steps_matches_reference_golden(golden_dir, tag):
    """TrajectorySampler.generate_model_rollouts, L = 5 with supplied eps, frames of every step + fp64 summed costs."""
    from robot_aware_control_b200 import DemoGoalState, State, TrajectorySampler

    gold, cfg, sd, actions, eps, states, masks = rollout_setup(golden_dir, tag)
    scene = np.load(os.path.join(golden_dir, "scene.npz"))
    N, L = int(gold["N"]), int(gold["L"])
    ts = TrajectorySampler(cfg, _model(cfg, sd))
    ts.set_noise(eps)
    start = State(img=scene["start_img"], state=np.array([0.3, 0.0, 0.2, 0.0, 0.0], dtype=np.float32), qpos=np.zeros(6))
    goal = DemoGoalState(imgs=list(scene["goal_imgs"]), masks=list(scene["goal_masks"]))
    r = ts.generate_model_rollouts(actions, start, goal, ret_obs=True, states=states, masks=masks)
    inv = np.empty(N, dtype=np.int64)
    inv[r["topk_idx"]] = np.arange(N)
    obs = r["obs"][inv]
    err = np.abs(obs - gold["obs"]).reshape(N, L, -1).max(2).max(0)
    print(f"g512 rollout {tag}: max-abs frame error per step {err}")
    if not tag.startswith("trained"):
        # north_star gate: identical random-init weights, actions and noise -> every frame within 1e-2
        assert err.max() < PIX_TOL, err
    else:
        # "trained-like" weights are NOT the north_star configuration; they are the harder case SURVEY 8(d) asks to be
        # looked at. The decoder head has 8x the gain, so the bf16 rounding of its 64-channel input alone is ~2e-3 on a
        # pixel, and the autoregressive loop amplifies ANY perturbation ~1.6x per step (measured on B200:
        # 0.0020 / 0.0046 / 0.010 / 0.016 / 0.024 vanilla, 0.0024 / 0.0045 / 0.0083 / 0.024 / 0.024 robot-aware).
        # Gate: the north_star tolerance for the first two predicted frames, 3e-2 after five.
        assert err[:2].max() < PIX_TOL and err.max() < 3e-2, err
    np.testing.assert_allclose(r["sum_cost"], gold["sum_cost"], rtol=3e-3)


@pytest.mark.parametrize("tag", ["l1", "config3"])
def test_train_step_g512_batch16(golden_dir, tag):
    """Two consecutive reference `_train_step`s at batch 16 / n_future 5 / g512: losses of both steps (the second one
    has gone through the CUDA backward + Adam), BatchNorm running statistics, and -- for the teacher-forced l1 config --
    the norms of the smooth-path (KL-only) gradients against the reference's own."""
    from robot_aware_control_b200 import SVGConvModel, SVGTrainer

    gold = np.load(os.path.join(golden_dir, f"train_g512_{tag}.npz"))
    kw = dict(lr=float(gold["lr"]), beta=float(gold["beta"]), beta1=0.9, n_future=5, n_past=1)
    if tag == "l1":
        cfg = cfg_for("l1", **kw)
    else:
        cfg = cfg_for("ra", model_use_future_mask=True, **kw)
    sd = so.make_state_dict(cfg, int(gold["weight_seed"]))
    model = SVGConvModel(cfg)
    model.load_state_dict(sd)
    model.train()
    trainer = SVGTrainer(cfg, model)
    batch, ep, eq = train_batch(int(gold["input_seed"]), cfg, tag != "l1")
    tokens = [True] + [False] * 4 if tag == "config3" else None
    keys = list(gold["keys"])
    for step in range(2):
        if tokens is not None:
            trainer.set_true_tokens(tokens)
        trainer.set_noise(ep, eq)
        losses = trainer.forward_backward(batch).cpu().numpy()
        print(f"train g512 {tag} step {step}: recon {losses[0]:.6f} (ref {float(gold[f'recon{step}']):.6f}) "
              f"kld {losses[1]:.6f} (ref {float(gold[f'kld{step}']):.6f})")
        np.testing.assert_allclose(losses[0], gold[f"recon{step}"], rtol=3e-3)
        np.testing.assert_allclose(losses[1], gold[f"kld{step}"], rtol=1e-2)
        if tag == "config3":
            np.testing.assert_allclose(losses[3], gold[f"world{step}"], rtol=1e-2)
        if tag == "l1" and step == 0:
            for i, k in enumerate(keys):
                if k.startswith("prior.") or k.startswith("prior_input_conv"):
                    gn = float(trainer.grad_of(k).double().norm())
                    assert abs(gn - gold["grad_norm0"][i]) <= 3e-2 * gold["grad_norm0"][i] + 1e-9, (k, gn, gold["grad_norm0"][i])
        trainer.optimizer_step()
        # post-Adam parameters: per-tensor norms against the reference's (lr 1e-4: a loose but global check that the
        # update has the reference's size everywhere)
        pn = np.array([float(dict(model.named_parameters())[k].double().norm()) for k in keys])
        np.testing.assert_allclose(pn, gold[f"param_norm{step}"], rtol=2e-4)
