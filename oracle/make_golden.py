"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on CPU with deterministic synthetic weights
(oracle.svg_oracle.make_state_dict) and injected noise. Run in the build container:

    python -m oracle.make_golden

The fixtures hold inputs seeds + reference outputs only (weights are regenerated from the seed), so they stay small.
They pin oracle/svg_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path on the GPU box.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, svg_oracle as so  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
G_DIM, Z_DIM = 128, 10


def inputs_forward(seed, B, cfg):
    g = torch.Generator().manual_seed(seed)
    d = {
        "image": torch.rand(2, B, 3, 48, 64, generator=g),
        "action": (torch.rand(2, B, cfg.action_dim, generator=g) - 0.5) * 0.1,
        "eps": torch.randn(2, B, cfg.z_dim, 6, 8, generator=g),
        "eps_post": torch.randn(2, B, cfg.z_dim, 6, 8, generator=g),
        "robot": torch.rand(3, B, cfg.robot_dim, generator=g),
        "mask": (torch.rand(3, B, 1, 48, 64, generator=g) > 0.8).float(),
    }
    return d


def synth_masks(seed, L, N):
    """per-candidate random rectangles, ~15-25 % coverage, float {0,1}."""
    rs = np.random.RandomState(seed)
    m = np.zeros((L + 1, N, 1, 48, 64), dtype=np.float32)
    for t in range(L + 1):
        for n in range(N):
            h, w = rs.randint(16, 28), rs.randint(20, 32)
            y, x = rs.randint(0, 48 - h), rs.randint(0, 64 - w)
            m[t, n, 0, y:y + h, x:x + w] = 1
    return torch.from_numpy(m)


class EpsFeeder:
    """Replaces GaussianConvLSTM.reparameterize (lstm.py:276-279): same arithmetic, noise popped from a queue."""

    def __init__(self):
        self.queue = []

    def __call__(self, module_self, mu, logvar):
        eps = self.queue.pop(0)
        return eps.mul(logvar.mul(0.5).exp()).add(mu)


class FakeRobotModel:
    def __init__(self, states, masks):
        self.states, self.masks = states, masks

    def predict_batch(self, data, thick=True):
        return self.states, self.masks


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = ref_shim.import_reference()
    dyn = mods["src.prediction.models.dynamics"]
    lstm_mod = sys.modules["src.prediction.models.lstm"]
    ts_mod = mods["src.cem.trajectory_sampler"]
    cem_mod = mods["src.cem.cem"]
    losses = mods["src.prediction.losses"]
    State, DemoGoalState = mods["src.utils.state"].State, mods["src.utils.state"].DemoGoalState
    feeder = EpsFeeder()
    lstm_mod.GaussianConvLSTM.reparameterize = lambda self, mu, logvar: feeder(self, mu, logvar)

    def build(cfg, seed):
        torch.manual_seed(1234)
        m = dyn.SVGConvModel(cfg)
        sd = so.make_state_dict(cfg, seed)
        missing = m.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        m.eval()
        return m

    # ------------------------------------------------------------------ G1/G2: SVGConvModel.forward
    for tag, kw in (("vanilla", dict(robot_aware=False)),
                    ("ra", dict(robot_aware=True, future_mask=True, future_robot_state=True))):
        cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, **kw)
        model = build(cfg, seed=11)
        B = 3
        d = inputs_forward(21, B, cfg)
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
                x_pred, skip, mu, logvar, mu_p, logvar_p = model.forward(d["image"][t], mask, robot, None, d["action"][t])
                out[f"x_pred{t}"] = x_pred.numpy()
                out[f"mu_p{t}"] = mu_p.numpy()
                out[f"logvar_p{t}"] = logvar_p.numpy()
                if t == 0:
                    out["h4"] = skip[3].numpy()
                    out["h1_sample"] = skip[0][:, :8].numpy()
            # posterior branch on a fresh hidden state (dynamics.py:613-629), prior noise first then posterior
            model.init_hidden(B)
            mask = robot = next_robot = None
            if cfg.model_use_mask:
                mask = torch.cat([d["mask"][0], d["mask"][1]], 1)
            if cfg.model_use_robot_state:
                robot = (d["robot"][0], d["robot"][1])
                next_robot = d["robot"][1]
            feeder.queue = [d["eps"][0].clone(), d["eps_post"][0].clone()]
            x_pred, _, mu, logvar, mu_p, logvar_p = model.forward(
                d["image"][0], mask, robot, None, d["action"][0], d["image"][1], mask, next_robot, None)
            out["post_x_pred"] = x_pred.numpy()
            out["post_mu"] = mu.numpy()
            out["post_logvar"] = logvar.numpy()
        np.savez_compressed(os.path.join(OUT, f"forward_{tag}.npz"), weight_seed=11, input_seed=21, B=B, **out)
        print("forward", tag, {k: v.shape for k, v in out.items()})

    # ------------------------------------------------------------------ G3/G4: generate_model_rollouts
    rs = np.random.RandomState(0)
    start_img = rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)
    goal_imgs = [rs.randint(0, 256, (48, 64, 3)).astype(np.uint8) for _ in range(2)]
    goal_mask = np.zeros((1, 48, 64), dtype=np.float32)
    goal_mask[:, :15] = 1  # widowx_VMPC_controller.py:338-340
    goal_masks = [goal_mask, goal_mask.copy()]
    np.savez_compressed(os.path.join(OUT, "scene.npz"), start_img=start_img, goal_imgs=np.stack(goal_imgs),
                        goal_masks=np.stack(goal_masks))
    N, L = 6, 3
    for tag, kw, extra in (("vanilla", dict(robot_aware=False), ()),
                           ("ra", dict(robot_aware=True, future_mask=True), ()),
                           ("ra_sparse", dict(robot_aware=True), ("--sparse_cost", "True"))):
        cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, extra=("--candidates_batch_size", str(N), "--topk", str(N)) + tuple(extra), **kw)
        model = build(cfg, seed=12)
        g = torch.Generator().manual_seed(31)
        actions = torch.cat([(torch.rand(N, L, 2, generator=g) - 0.5) * 0.1, torch.zeros(N, L, 3)], 2)
        eps = torch.randn(L, N, cfg.z_dim, 6, 8, generator=g)
        states = torch.rand(L + 1, N, 5, generator=g)
        masks = synth_masks(41, L, N)
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
        obs = r["obs"][inv]  # back to candidate order
        np.savez_compressed(os.path.join(OUT, f"rollout_{tag}.npz"), weight_seed=12, input_seed=31, mask_seed=41, N=N,
                            L=L, sum_cost=r["sum_cost"], obs=obs.astype(np.float32))
        print("rollout", tag, r["sum_cost"])

    # ------------------------------------------------------------------ G5: CEMPolicy.get_action
    cfg = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, robot_aware=False, extra=("--candidates_batch_size", "8"))
    model = build(cfg, seed=13)
    I, N, L, K = 3, 8, 3, 3
    g = torch.Generator().manual_seed(51)
    noise = torch.randn(I, N, L, 2, generator=g)
    eps = torch.randn(I, L, N, cfg.z_dim, 6, 8, generator=g)

    class FakeNormal:
        calls = 0

        def __init__(self, mean, std):
            self.mean, self.std = mean, std

        def sample(self, shape):
            i = FakeNormal.calls
            FakeNormal.calls += 1
            return self.mean + self.std * noise[i]

    cem_mod.Normal = FakeNormal
    cem_mod.trange = lambda n, desc=None: range(n)
    policy = cem_mod.CEMPolicy(cfg, model, horizon=L + 1, opt_iter=I, action_candidates=N, topk=K, init_std=0.03)
    feeder.queue = [eps[i, t].clone() for i in range(I) for t in range(L)]
    record = []
    orig = policy.traj_sampler.generate_model_rollouts

    def wrapped(*a, **k):
        r = orig(*a, **k)
        record.append(np.array(r["sum_cost"], dtype=np.float64).copy())
        return r

    policy.traj_sampler.generate_model_rollouts = wrapped
    start = State(img=start_img)
    goal = DemoGoalState(imgs=goal_imgs, masks=goal_masks)
    mean = policy.get_action(start, goal, 0, 0)
    np.savez_compressed(os.path.join(OUT, "cem_vanilla.npz"), weight_seed=13, input_seed=51, I=I, N=N, L=L, K=K,
                        init_std=0.03, mean=mean, sum_costs=np.stack(record))
    print("cem mean", mean)

    # ------------------------------------------------------------------ G6: costs, criteria, top-k
    g = torch.Generator().manual_seed(61)
    B = 5
    curr = torch.rand(B, 3, 48, 64, generator=g)
    goal_t = torch.rand(3, 48, 64, generator=g)
    cmask = (torch.rand(B, 1, 48, 64, generator=g) > 0.7).float()
    gmask = (torch.rand(1, 48, 64, generator=g) > 0.7).float()
    cfg_l2 = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM)
    cfg_dc = ref_shim.make_cfg(g_dim=G_DIM, z_dim=Z_DIM, robot_aware=True)
    l2 = losses.RobotWorldCost(cfg_l2)(State(img=curr.clone()), State(img=goal_t.clone()))
    dc = losses.RobotWorldCost(cfg_dc)(State(img=curr.clone(), mask=cmask), State(img=goal_t.clone(), mask=gmask))
    pred = torch.rand(B, 3, 48, 64, generator=g)
    mu1, lv1, mu2, lv2 = (torch.randn(B, Z_DIM, 6, 8, generator=g) * 0.5 for _ in range(4))
    out = {
        "img_l2": np.asarray(l2), "img_dontcare": np.asarray(dc),
        "l1": losses.l1_criterion(pred.clone(), curr.clone()).numpy(),
        "dontcare_l1_w0": losses.dontcare_l1_criterion(pred.clone(), curr.clone(), cmask, 0.0).numpy(),
        "dontcare_l1_w05": losses.dontcare_l1_criterion(pred.clone(), curr.clone(), cmask, 0.5).numpy(),
        "kl": losses.kl_criterion(mu1, lv1, mu2, lv2, B).numpy(),
        "robot_mse": losses.robot_mse_criterion(pred.clone(), curr.clone(), cmask).numpy(),
        "world_mse": losses.world_mse_criterion(pred.clone(), curr.clone(), cmask).numpy(),
    }
    # top-k: torch.topk (cem.py:97) on vectors without ties at the K boundary; torch.sort(stable) for the tie rule
    rs = np.random.RandomState(7)
    for name, n, k in (("a", 2000, 200), ("b", 16384, 1638), ("c", 100, 5), ("d", 64, 64)):
        c = rs.randn(n) * 3 - 20000
        v, idx = torch.from_numpy(c).topk(k)
        out[f"topk_{name}_costs"] = c
        out[f"topk_{name}_idx"] = idx.numpy()
    ties = np.round(rs.randn(512) * 2).astype(np.float64)
    st = torch.sort(torch.from_numpy(ties), descending=True, stable=True)[1][:37]
    out["topk_ties_costs"] = ties
    out["topk_ties_idx"] = st.numpy()
    elite_act = torch.randn(200, 4, 2, generator=g) * 0.02
    std, mean_ = torch.std_mean(elite_act, dim=0)
    out["refit_act"] = elite_act.numpy()
    out["refit_mean"] = mean_.numpy()
    out["refit_std"] = torch.max(0.001 * torch.ones_like(std), std).numpy()
    np.savez_compressed(os.path.join(OUT, "costs.npz"), input_seed=61, **out)
    print("costs", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
