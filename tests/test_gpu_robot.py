"""GPU tests of the device-side robot state / mask producer (csrc/robot_kernels.cu, robot.py) and of the robot-aware
plan that uses it without a host round trip:

* rac_predict_states == the states the UNMODIFIED reference returned (tests/golden/robot_states.npz), bit for bit
* rac_render_masks == oracle/robot_oracle.py::render_masks except on rounding-sensitive silhouette pixels (this half
  has no reference pin: the reference renders MuJoCo meshes)
* CEMPolicy.get_action with a DeviceRobotModel is ONE rac_cem_plan call and equals the per-iteration path bit for bit
* the remaining reconstruction criteria (mse, dontcare_mse, movement-weighted l1 kinds) == the reference's values"""
import os

import numpy as np
import pytest
import torch

from oracle import robot_oracle as ro
from oracle import svg_oracle as so
from oracle.make_golden_robot import PUSH_HEIGHT, inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["wx250s", "franka"])
def test_predict_states_bit_equal_to_reference_golden(golden_dir, kind):
    from robot_aware_control_b200 import DeviceRobotModel
    from robot_aware_control_b200.robot import normalized_start_state

    gold = np.load(os.path.join(golden_dir, "robot_states.npz"))
    rm = DeviceRobotModel(kind=kind, push_height=PUSH_HEIGHT)
    low = torch.from_numpy(ro.LOW)[None]
    high = torch.from_numpy(ro.HIGH)[None]
    for seed in (1, 2):
        actions, start = inputs(seed)
        sn = normalized_start_state(start, "control_" + kind, low, high)
        np.testing.assert_array_equal(sn.numpy(), gold[f"{kind}_start_norm_{seed}"])
        st = rm.predict_states(sn.cuda().contiguous(), actions.cuda())
        np.testing.assert_array_equal(st.cpu().numpy(), gold[f"{kind}_states_{seed}"])
        # the reference robot-model interface: predict_batch(data) as trajectory_sampler.py:93-107 builds it
        N, L = actions.shape[0], actions.shape[1]
        states = torch.zeros(L + 1, N, 5)
        states[0, :] = sn
        data = {"states": states, "qpos": torch.zeros(L + 1, N, 6), "actions": actions.permute(1, 0, 2),
                "low": low.repeat(N, 1), "high": high.repeat(N, 1)}
        s2, m2 = rm.predict_batch(data, thick=True)
        np.testing.assert_array_equal(s2.cpu().numpy(), gold[f"{kind}_states_{seed}"])
        assert m2.shape == (L + 1, N, 1, 48, 64) and set(np.unique(m2.cpu().numpy())) <= {0.0, 1.0}


def test_render_masks_match_capsule_oracle():
    from robot_aware_control_b200 import DeviceRobotModel

    rm = DeviceRobotModel(kind="wx250s", push_height=0.1)
    g = torch.Generator().manual_seed(3)
    raw = torch.zeros(3, 4, 5)
    raw[..., 0] = 0.2 + 0.3 * torch.rand(3, 4, generator=g)
    raw[..., 1] = -0.2 + 0.4 * torch.rand(3, 4, generator=g)
    raw[..., 2] = 0.1
    raw[..., :2] += torch.from_numpy(ro.LOCO_WX250S_DIFF).float()
    states = ((raw - torch.from_numpy(ro.LOW)) / torch.from_numpy(ro.HIGH - ro.LOW)).contiguous()
    m = rm.c_model
    for thick in (False, True):
        got = rm.render(states.cuda(), thick=thick).cpu().numpy()
        ref, edge = ro.render_masks(states.numpy(), "wx250s", list(m.cam_center), list(m.cam_minv), m.shoulder_z, m.l_upper,
                                    m.l_fore, m.l_wrist, m.pitch, list(m.radius), rm.thick_extra if thick else 0.0,
                                    margin=2e-4)
        diff = got != ref
        assert not (diff & ~edge).any(), int((diff & ~edge).sum())   # only pixels within 0.2 mm of a silhouette may differ
        assert diff.mean() < 2e-3
        cover = got.mean((2, 3, 4))
        assert (cover > 0.01).all() and (cover < 0.6).all(), cover


def test_robot_aware_plan_with_device_robot_model_is_one_call(golden_dir):
    """get_action with a DeviceRobotModel: the fused rac_cem_plan (states + masks predicted per iteration on the
    device) equals the per-iteration path that calls predict_states / render from Python, bit for bit."""
    from robot_aware_control_b200 import CEMPolicy, DemoGoalState, DeviceRobotModel, State, SVGConvModel

    scene = np.load(os.path.join(golden_dir, "scene.npz"))
    cfg = so.make_cfg(g_dim=128, z_dim=10, model_use_mask=True, model_use_future_mask=True, model_use_robot_state=True,
                      reconstruction_loss="dontcare_l1", reward_type="dontcare", experiment="control_wx250s")
    model = SVGConvModel(cfg)
    model.load_state_dict(so.make_state_dict(cfg, 4))
    model.eval()
    I, N, L, K = 3, 24, 3, 5
    noise = torch.randn(I, N, L, 2, generator=torch.Generator().manual_seed(8))
    start = State(img=scene["start_img"], state=np.array([0.3, 0.02, 0.1, 0.0, 0.0], dtype=np.float32), qpos=np.zeros(6))
    goal = DemoGoalState(imgs=list(scene["goal_imgs"]), masks=list(scene["goal_masks"]))
    rm = DeviceRobotModel(cfg, kind="wx250s", push_height=0.1)
    outs = []
    for mode in ("fused", "stepwise"):
        robot = rm if mode == "fused" else DeviceRobotModel(cfg, kind="wx250s", push_height=0.1,
                                                            mask_fn=lambda s: rm.render(s, thick=True))
        pol = CEMPolicy(cfg, model, horizon=L + 1, opt_iter=I, action_candidates=N, topk=K, init_std=0.03,
                        robot_model=robot)
        pol._seed = 77
        pol.set_noise(noise)
        launches = model.launch_count()
        mean = pol.get_action(start, goal, 0, 0)
        outs.append((mean, pol.last_costs.cpu().numpy(), pol.last_elite_idx.cpu().numpy()))
        assert model.launch_count() > launches
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)
    assert np.all(np.isfinite(outs[0][1])) and np.all(outs[0][1] < 0)
    # without start.state the robot model cannot run
    with pytest.raises(ValueError):
        CEMPolicy(cfg, model, horizon=L + 1, opt_iter=1, action_candidates=N, topk=K, robot_model=rm).get_action(
            State(img=scene["start_img"]), goal, 0, 0)


def test_remaining_recon_criteria_match_reference_arithmetic():
    """mse_criterion / dontcare_mse_criterion / batch-weighted l1 kinds (losses.py:11-50) against the oracle restatement
    (pinned to the reference through the train_vanilla_mse / train_ra_dcmse / train_ra_bw goldens)."""
    from robot_aware_control_b200 import dontcare_l1_criterion, dontcare_mse_criterion, l1_criterion, mse_criterion

    g = torch.Generator().manual_seed(5)
    B = 6
    pred, tgt = torch.rand(B, 3, 48, 64, generator=g), torch.rand(B, 3, 48, 64, generator=g)
    mask = (torch.rand(B, 1, 48, 64, generator=g) > 0.7).float()
    bw = torch.tensor([3.0, 1.0, 1.0, 3.0, 3.0, 1.0])
    np.testing.assert_allclose(mse_criterion(pred, tgt).item(), so.mse_criterion(pred, tgt).item(), rtol=1e-5)
    np.testing.assert_allclose(dontcare_mse_criterion(pred, tgt, mask, 0.25).item(),
                               so.dontcare_mse_criterion(pred, tgt, mask, 0.25).item(), rtol=1e-5)
    np.testing.assert_allclose(l1_criterion(pred, tgt, bw).item(), so.l1_criterion(pred, tgt, bw).item(), rtol=1e-5)
    np.testing.assert_allclose(dontcare_l1_criterion(pred, tgt, mask, 0.5, bw).item(),
                               so.dontcare_l1_criterion(pred, tgt, mask, 0.5, bw).item(), rtol=1e-5)
