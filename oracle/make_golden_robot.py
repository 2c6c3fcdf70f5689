"""TEST INFRASTRUCTURE -- golden vectors of the robot STATE producer: runs the UNMODIFIED reference
`TrajectorySampler.generate_model_rollouts` (src/cem/trajectory_sampler.py:86-109: frame shift + normalisation of the
start state) with the UNMODIFIED `WX250sAnalyticalModel.predict_batch` / `FrankaAnalyticalModel.predict_batch`
(src/dataset/wx250s/wx250s_model.py:57-182, src/dataset/franka/franka_model.py:30-95) and records the states they
return. Only the parts that need the lab setup are stubbed: the Interbotix IK call `bot.arm.set_ee_pose_components`
(its result, qpos, never enters the states), the ROS IK service of the Franka model, and the MuJoCo mask render
`env.generate_masks` (returns empty masks -- the mask half is not pinned, see robot_aware_control_b200/robot.py).

    python -m oracle.make_golden_robot   ->  tests/golden/robot_states.npz
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, svg_oracle as so  # noqa: E402
from oracle.make_golden import EpsFeeder  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N, L = 7, 5
PUSH_HEIGHT = 0.1234


def inputs(seed):
    g = torch.Generator().manual_seed(seed)
    actions = torch.cat([(torch.rand(N, L, 2, generator=g) - 0.5) * 0.1, torch.zeros(N, L, 3)], 2)
    start_state = np.array([0.31 + 0.01 * seed, -0.07, 0.2, 0.3, 0.9], dtype=np.float32)
    return actions, start_state


class _Env:
    def generate_masks(self, qpos):
        return [np.zeros((48, 64), dtype=np.uint8) for _ in range(len(qpos))]


def main():
    import importlib

    mods = ref_shim.import_reference()
    ts_mod = mods["src.cem.trajectory_sampler"]
    lstm_mod = sys.modules["src.prediction.models.lstm"]
    dyn = mods["src.prediction.models.dynamics"]
    losses = mods["src.prediction.losses"]
    State, DemoGoalState = mods["src.utils.state"].State, mods["src.utils.state"].DemoGoalState
    wx_mod = importlib.import_module("src.dataset.wx250s.wx250s_model")
    fr_mod = importlib.import_module("src.dataset.franka.franka_model")
    feeder = EpsFeeder()
    lstm_mod.GaussianConvLSTM.reparameterize = lambda self, mu, logvar: feeder(self, mu, logvar)
    scene = np.load(os.path.join(OUT, "scene.npz"))
    out = {}
    for tag, experiment in (("wx250s", "control_wx250s"), ("franka", "control_franka")):
        cfg = ref_shim.make_cfg(g_dim=128, z_dim=10, robot_aware=True,
                                extra=("--candidates_batch_size", str(N), "--topk", str(N), "--experiment", experiment,
                                       "--robot_joint_dim", "6" if tag == "wx250s" else "7"))
        torch.manual_seed(0)
        model = dyn.SVGConvModel(cfg)
        model.load_state_dict(so.make_state_dict(cfg, 3))
        model.eval()
        if tag == "wx250s":
            rm = wx_mod.WX250sAnalyticalModel.__new__(wx_mod.WX250sAnalyticalModel)
            rm.bot = SimpleNamespace(arm=SimpleNamespace(set_ee_pose_components=lambda **kw: (np.zeros(6), True)))
            rm.push_height, rm.default_pitch, rm.default_roll = PUSH_HEIGHT, 1.5, 0.0
        else:
            rm = fr_mod.FrankaAnalyticalModel.__new__(fr_mod.FrankaAnalyticalModel)

            def send_ik_request(q0, waypoints):
                n, t = waypoints.shape[0], waypoints.shape[1]
                return SimpleNamespace(num_traj=n, traj_length=t, joint_dim=7, joint_angles=np.zeros(n * t * 7))

            rm.ik_solver = SimpleNamespace(send_ik_request=send_ik_request)
        rm._config = cfg
        rm.env = rm.env_thick = _Env()
        rm._img_transform = lambda m: torch.zeros(1, 48, 64)
        recorded = {}
        orig = rm.predict_batch

        def wrapped(data, thick=False, _orig=orig, _rec=recorded):
            _rec["in_states0"] = data["states"][0, 0].clone().numpy()
            s, m = _orig(data, thick=thick)
            _rec["states"] = s.clone().numpy()
            return s, m

        rm.predict_batch = wrapped
        sampler = ts_mod.TrajectorySampler.__new__(ts_mod.TrajectorySampler)
        sampler.cfg, sampler.model, sampler.cost = cfg, model, losses.RobotWorldCost(cfg)
        sampler.low = torch.from_numpy(np.array([[0.015, -0.3, 0.1, 0, 0]], dtype=np.float32))
        sampler.high = torch.from_numpy(np.array([[0.55, 0.3, 0.4, 1, 1]], dtype=np.float32))
        sampler.robot_model = rm
        for seed in (1, 2):
            actions, start_state = inputs(seed)
            feeder.queue = [torch.zeros(N, cfg.z_dim, 6, 8) for _ in range(L)]
            start = State(img=scene["start_img"], state=start_state.copy(), qpos=np.zeros(6 if tag == "wx250s" else 7))
            goal = DemoGoalState(imgs=list(scene["goal_imgs"]), masks=list(scene["goal_masks"]))
            sampler.generate_model_rollouts(actions, start, goal)
            out[f"{tag}_states_{seed}"] = recorded["states"].astype(np.float32)
            out[f"{tag}_start_norm_{seed}"] = recorded["in_states0"].astype(np.float32)
            print(tag, seed, recorded["states"][:, 0])
    np.savez_compressed(os.path.join(OUT, "robot_states.npz"), N=N, L=L, push_height=PUSH_HEIGHT, **out)


if __name__ == "__main__":
    main()
