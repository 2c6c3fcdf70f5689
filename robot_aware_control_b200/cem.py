"""Drop-ins for the reference planner: `TrajectorySampler` (src/cem/trajectory_sampler.py:15-199) and `CEMPolicy`
(src/cem/cem.py:14-111), same constructors / method signatures / return types, on B200.

What changes underneath: the autoregressive SVG rollout, compositing, robot-pixel zeroing and the RobotWorldCost run
as one stream of CUDA kernels per call (rac_rollout_cost) with a single device->host read of the fp64 summed costs;
sampling, elite top-k and the Gaussian refit run on the device as well. When no host-side robot model is involved the
whole `get_action` is one `rac_cem_plan` call (all iterations device resident, one read of the final mean).

Multi-GPU (new, the reference is single-process): pass `process_group`; rank r rolls out candidates
[r*N/R, (r+1)*N/R), only the per-candidate costs are all-gathered (NCCL), the elite refit is replicated.
"""
import ctypes as C
from collections import defaultdict

import numpy as np
import torch

from . import _lib, parallel
from .losses import RobotWorldCost
from .robot import DeviceRobotModel, normalized_start_state
from .state import DemoGoalState, State

# frame offsets between robots (reference src/utils/camera_calibration.py:176-177)
LOCO_FRANKA_DIFF = np.array([-0.365, -0.06103333])
LOCO_WX250S_DIFF = np.array([-0.13, -0.01])


def _needs_robot(cfg):
    return bool(getattr(cfg, "model_use_robot_state", False) or getattr(cfg, "model_use_mask", False)
                or getattr(cfg, "black_robot_input", False) or "dontcare" in getattr(cfg, "reward_type", ""))


def _zero_robot(cfg):
    return bool("dontcare" in getattr(cfg, "reconstruction_loss", "") or getattr(cfg, "black_robot_input", False))


class TrajectorySampler(object):
    def __init__(self, cfg, model, cam_ext=None, franka_ik=None, wx250s_bot=None, push_height=None,
                 default_pitch=None, default_roll=None, robot_model=None) -> None:
        """`robot_model`: object with `predict_batch(data, thick=True) -> (states (T+1,N,5), masks (T+1,N,1,H,W))`
        (reference: WX250sAnalyticalModel / FrankaAnalyticalModel, MuJoCo + IK, built inside the reference constructor
        at trajectory_sampler.py:26-33; here it is injected because the simulator is outside the hot path)."""
        super().__init__()
        self.cfg = cfg
        self.model = model
        self.cost = RobotWorldCost(cfg)
        self.low = torch.from_numpy(np.array([0.015, -0.3, 0.1, 0, 0], dtype=np.float32)).unsqueeze_(0)
        self.high = torch.from_numpy(np.array([0.55, 0.3, 0.4, 1, 1], dtype=np.float32)).unsqueeze_(0)
        self.robot_model = robot_model
        self._lib = _lib.load()
        self._eps = None
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._noise_ctr = 0
        self.cand_offset = 0

    # ------------------------------------------------------------------ helpers
    def set_noise(self, eps):
        """Test hook: (L, N, z_dim, 6, 8) prior noise for the next rollout instead of Philox."""
        self._eps = eps

    def _goal_tensors(self, goal):
        dev = self.model._device
        imgs = np.stack([np.ascontiguousarray(g) for g in goal.imgs]).astype(np.uint8)
        goal_imgs = torch.from_numpy(imgs).to(dev, non_blocking=True)
        goal_masks = None
        if goal.masks is not None:
            gm = np.stack([np.asarray(g, dtype=np.float32).reshape(48, 64) for g in goal.masks])
            goal_masks = torch.from_numpy(gm).to(dev, non_blocking=True)
        return goal_imgs, goal_masks

    def _robot_inputs(self, action_sequences, start, N, T):
        """trajectory_sampler.py:86-109."""
        cfg = self.cfg
        if self.robot_model is None:
            raise ValueError("this configuration needs robot states/masks (model_use_robot_state / model_use_mask / "
                             "dontcare): pass robot_model= to TrajectorySampler or states=/masks= to the call")
        states = torch.zeros((T + 1, N, 5), dtype=torch.float32)
        qpos = torch.zeros((T + 1, N, getattr(cfg, "robot_joint_dim", 6)), dtype=torch.float32)
        start_state = torch.tensor(np.asarray(start.state, dtype=np.float32))
        exp = getattr(cfg, "experiment", "")
        if exp == "control_franka":
            # (float32 tensor + float64 numpy constant: summed in double, stored back as float32, as the reference)
            start_state[:2] = start_state[:2] + torch.from_numpy(LOCO_FRANKA_DIFF)
        elif exp == "control_wx250s":
            start_state[:2] = start_state[:2] + torch.from_numpy(LOCO_WX250S_DIFF)
        states[0, :] = (start_state - self.low) / (self.high - self.low)
        qpos[0, :] = torch.tensor(np.asarray(start.qpos, dtype=np.float32))
        data = {"states": states, "qpos": qpos, "actions": action_sequences.permute(1, 0, 2),
                "low": self.low.repeat(N, 1), "high": self.high.repeat(N, 1)}
        return self.robot_model.predict_batch(data, thick=True)

    def _rollout_device(self, actions_dev, start_img_dev, goal_imgs, goal_masks, states, masks, eps, n, steps,
                        sum_cost, obs_out=None, step_cost=None, cand_offset=0, noise_ctr=None, peer=None):
        """One rac_rollout_cost call on device tensors (no host traffic). `peer` = (device pointer of the array of
        per-rank gathered-cost buffers, world size, element offset of this shard): the cost kernel of the last step
        then stores the finished costs into every rank's buffer over NVLink peer memory."""
        cfg = self.cfg
        m = self.model
        m.prepare(n)
        r = _lib.RacRollout()
        r.n, r.steps, r.cand_offset = n, steps, cand_offset
        r.actions = _lib.ptr(actions_dev)
        r.start_img = _lib.ptr(start_img_dev)
        r.goal_imgs = _lib.ptr(goal_imgs)
        r.num_goals = goal_imgs.shape[0]
        r.goal_masks = _lib.ptr(goal_masks)
        if states is not None:
            r.states = _lib.ptr(states)
            r.state_t_stride = states.stride(0)
        if masks is not None:
            r.masks = _lib.ptr(masks)
            r.mask_t_stride = masks.stride(0)
        r.eps = _lib.ptr(eps)
        r.seed = self._seed
        r.noise_ctr_base = self._noise_ctr if noise_ctr is None else noise_ctr
        r.sample_mean = int(bool(getattr(cfg, "sample_mean", False)))
        r.zero_robot = int(_zero_robot(cfg))
        r.dontcare_cost = int(getattr(cfg, "reward_type", "") == "dontcare")
        r.sparse_cost = int(bool(getattr(cfg, "sparse_cost", False)))
        r.world_cost_weight = float(getattr(cfg, "world_cost_weight", 1.0))
        r.obs_out = _lib.ptr(obs_out)
        r.step_cost_out = _lib.ptr(step_cost)
        r.sum_cost = _lib.ptr(sum_cost)
        if peer is not None:
            r.peer_cost_bufs, r.peer_world, r.peer_offset = C.c_void_p(int(peer[0])), int(peer[1]), int(peer[2])
        # robot_cost_weight: RobotL2Cost contributes 0.0 inside CEM because State.state is None (losses.py:189-190)
        _lib.check(self._lib.rac_rollout_cost(m.handle, C.byref(r), _lib.stream_ptr()), m.handle, "rac_rollout_cost")
        if noise_ctr is None:
            self._noise_ctr += steps

    # ------------------------------------------------------------------ reference interface
    @torch.no_grad()
    def generate_model_rollouts(self, action_sequences, start: State, goal: DemoGoalState, opt_traj=None,
                                ret_obs=False, ret_step_cost=False, suppress_print=True, states=None, masks=None,
                                opt_robot=None):
        """trajectory_sampler.py:35-199. Extra keywords `states` / `masks` feed precomputed robot states
        (T+1, N, 5) and masks (T+1, N, 1, H, W) instead of calling robot_model.predict_batch; `opt_robot` =
        (states (T+1, 1, 5), masks (T+1, 1, 1, H, W)) of the expert trajectory `opt_traj` when no robot_model exists."""
        cfg = self.cfg
        m = self.model
        dev = m._device
        action_sequences = torch.as_tensor(action_sequences, dtype=torch.float32)
        N = len(action_sequences)
        T = action_sequences.shape[1]
        if m.training:
            raise NotImplementedError("planning needs an eval-mode model: call model.eval()")
        goal_imgs, goal_masks = self._goal_tensors(goal)
        start_img = torch.from_numpy(np.ascontiguousarray(start.img).astype(np.uint8)).to(dev, non_blocking=True)
        if _needs_robot(cfg) and (states is None or masks is None):
            states, masks = self._robot_inputs(action_sequences.cpu(), start, N, T)
        if states is not None:
            states = states.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        if masks is not None:
            masks = masks.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        need_masks = bool(getattr(cfg, "model_use_mask", False)) or _zero_robot(cfg) or getattr(cfg, "reward_type", "") == "dontcare"
        if not need_masks:
            masks = None
        if not getattr(cfg, "model_use_robot_state", False):
            states = None
        actions = action_sequences.to(dev, non_blocking=True).contiguous()
        sum_cost = torch.empty(N, dtype=torch.float64, device=dev)
        obs = torch.empty(T, N, 48, 64, 4, device=dev) if ret_obs else None
        step_cost = torch.empty(T, N, device=dev) if ret_step_cost else None
        eps = None
        if self._eps is not None:
            eps = self._eps.to(dev, dtype=torch.float32).contiguous()
            self._eps = None
        self._rollout_device(actions, start_img, goal_imgs, goal_masks, states, masks, eps, N, T, sum_cost, obs,
                             step_cost, cand_offset=self.cand_offset)
        rollouts = defaultdict(float)
        if opt_traj is not None:
            # the expert trajectory is candidate N of the reference's batch (:62-68,183-187); here it is rolled out
            # in a second pass of the same batch size (row 0) so that the workspace is not re-sized
            ot = torch.as_tensor(opt_traj, dtype=torch.float32)
            ot = torch.cat([ot, torch.zeros((len(ot), actions.shape[2] - ot.shape[1]))], 1).to(dev)
            acts2 = actions.clone()
            acts2[0] = ot
            states2, masks2 = states, masks
            if _needs_robot(cfg):
                # the reference appends the expert trajectory BEFORE robot_model.predict_batch (:62-68,101-109): its
                # robot states / masks are those predicted for the expert's own actions, not candidate 0's
                if opt_robot is not None:
                    s_ot, m_ot = opt_robot
                elif self.robot_model is not None:
                    s_ot, m_ot = self._robot_inputs(ot[None].cpu(), start, 1, T)
                else:
                    raise ValueError("opt_traj with a robot-aware configuration needs the expert trajectory's robot "
                                     "states / masks: pass opt_robot=(states (T+1,1,5), masks (T+1,1,1,H,W)) or a robot_model")
                if states is not None:
                    states2 = states.clone()
                    states2[:, 0] = s_ot.to(dev, dtype=torch.float32)[:, 0]
                if masks is not None:
                    masks2 = masks.clone()
                    masks2[:, 0] = m_ot.to(dev, dtype=torch.float32)[:, 0]
            sc2 = torch.empty(N, dtype=torch.float64, device=dev)
            obs2 = torch.empty(T, N, 48, 64, 4, device=dev)
            self._rollout_device(acts2, start_img, goal_imgs, goal_masks, states2, masks2, None, N, T, sc2, obs2, None,
                                 cand_offset=self.cand_offset)
            rollouts["optimal_sum_cost"] = float(sc2[0].item())
            rollouts["optimal_obs"] = obs2[:, 0, :, :, :3].permute(0, 3, 1, 2).cpu().numpy()
        sum_cost_np = sum_cost.cpu().numpy()  # the one device->host read of the rollout
        rollouts["sum_cost"] = sum_cost_np
        if ret_obs:
            topk_idx = np.argsort(sum_cost_np)[-getattr(cfg, "topk", 5):]
            sel = obs[:, torch.from_numpy(topk_idx).to(dev)]  # (T, K, H, W, 4)
            rollouts["topk_idx"] = topk_idx
            rollouts["obs"] = sel[..., :3].permute(1, 0, 4, 2, 3).cpu().numpy()  # (K, T, 3, H, W)
        if ret_step_cost:
            rollouts["step_cost"] = step_cost.transpose(0, 1).cpu().numpy()
        return rollouts


class CEMPolicy(object):
    """Given the current state and goal images, use CEM to find the best actions (reference cem.py:14)."""

    def __init__(self, cfg, model, horizon=5, opt_iter=10, action_candidates=100, topk=5, init_std=1.0, cam_ext=None,
                 franka_ik=None, wx250s_bot=None, push_height=None, default_pitch=None, default_roll=None,
                 robot_model=None, process_group=None, noise_source="torch", verbose=False):
        self.horizon = horizon
        self.optimization_iter = opt_iter
        self.num_actions = action_candidates
        self.K = topk
        self.init_std = init_std
        self.sparse_cost = getattr(cfg, "sparse_cost", False)
        self.action_dim = 2
        self.cfg = cfg
        self.model = model
        self.traj_sampler = TrajectorySampler(cfg, self.model, cam_ext=cam_ext, franka_ik=franka_ik,
                                              wx250s_bot=wx250s_bot, push_height=push_height,
                                              default_pitch=default_pitch, default_roll=default_roll,
                                              robot_model=robot_model)
        self.plot_rollouts = getattr(cfg, "debug_cem", False)
        self.process_group = process_group
        if noise_source not in ("torch", "philox"):
            raise ValueError("noise_source must be 'torch' (reference CPU generator stream) or 'philox' (on device)")
        self.noise_source = noise_source
        self.verbose = verbose
        self._lib = _lib.load()
        self._noise = None
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._plans = 0
        self.last_costs = None
        self.last_elite_idx = None
        self.last_std = None
        self.precomputed_robot = None  # optional (states, masks) device tensors reused for every iteration

    def set_noise(self, noise):
        """Test hook: (opt_iter, N, horizon-1, 2) standard normals for the next get_action."""
        self._noise = noise

    def _draw_noise(self, I, N, L, dev):
        if self._noise is not None:
            nz, self._noise = self._noise, None
            return nz.to(dev, dtype=torch.float32).contiguous()
        if self.noise_source == "torch":
            # same stream of the global CPU generator as I x Normal(mean, std).sample((N,)) in the reference (cem.py:80-81)
            return torch.randn(I, N, L, 2).pin_memory().to(dev, non_blocking=True)
        return None

    @torch.no_grad()
    def get_action(self, start, goal, ep_num, step, opt_traj=None):
        """cem.py:56-111. Returns the refit mean, np.float32 (horizon-1, 2)."""
        self.ep_num, self.step = ep_num, step
        m = self.model
        dev = m._device
        if m.training:
            raise NotImplementedError("planning needs an eval-mode model: call model.eval()")
        noise = self._draw_noise(self.optimization_iter, self.num_actions, self.horizon - 1, dev)
        goal_imgs, goal_masks = self.traj_sampler._goal_tensors(goal)
        start_img = torch.from_numpy(np.ascontiguousarray(start.img).astype(np.uint8)).to(dev, non_blocking=True)
        mean = self.plan_device(start_img, goal_imgs, goal_masks, noise, start=start, goal=goal, opt_traj=opt_traj)
        out = mean.cpu().numpy()  # the one device->host read of the plan
        if self.verbose:
            print("Mean actions:", out)
        return out

    @torch.no_grad()
    def plan_device(self, start_img, goal_imgs, goal_masks=None, noise=None, start=None, goal=None, opt_traj=None):
        """The plan on device-resident inputs: start_img uint8 (H,W,3), goal_imgs uint8 (G,H,W,3), goal_masks fp32
        (G,H,W) or None, noise fp32 (opt_iter, N, horizon-1, 2) or None (Philox). Returns the mean as a (horizon-1, 2)
        CUDA tensor without synchronising. `start` / `goal` are only needed when a host-side robot model runs."""
        cfg = self.cfg
        T, N, I, K = self.horizon, self.num_actions, self.optimization_iter, self.K
        L = T - 1
        m = self.model
        dev = m._device
        ts = self.traj_sampler
        world, rank = parallel.world_info(self.process_group)
        lo, hi = parallel.shard_range(N, rank, world)
        n_local = hi - lo
        # a DeviceRobotModel predicts states (and masks) on the GPU: no host round trip per iteration
        dev_robot = ts.robot_model if isinstance(ts.robot_model, DeviceRobotModel) else None
        if self.precomputed_robot is not None or not _needs_robot(cfg):
            dev_robot = None
        host_robot = _needs_robot(cfg) and self.precomputed_robot is None and dev_robot is None
        start_norm = None
        if dev_robot is not None:
            if start is None or start.state is None:
                raise ValueError("a robot model needs start.state (the current end-effector state)")
            start_norm = normalized_start_state(start.state, getattr(cfg, "experiment", ""), ts.low, ts.high)
            start_norm = start_norm.to(dev, dtype=torch.float32).contiguous()
        fused_robot = dev_robot is None or dev_robot.mask_fn is None
        A = m._c.action_dim
        plan_seed = (self._seed + 0x9E3779B97F4A7C15 * (self._plans + 1)) & 0xFFFFFFFFFFFFFFFF
        self._plans += 1
        ts._seed = plan_seed

        if world == 1 and not host_robot and fused_robot and opt_traj is None and not self.plot_rollouts:
            # ---- whole plan on the device: one C call
            m.prepare(N)
            c = _lib.RacCem()
            c.n, c.steps, c.iters, c.topk = N, L, I, K
            c.init_std, c.clamp, c.std_floor = float(self.init_std), 0.05, 0.001
            c.noise = _lib.ptr(noise)
            r = c.rollout
            r.start_img, r.goal_imgs, r.num_goals = _lib.ptr(start_img), _lib.ptr(goal_imgs), goal_imgs.shape[0]
            r.goal_masks = _lib.ptr(goal_masks)
            if self.precomputed_robot is not None:
                states, masks = self.precomputed_robot
                if getattr(cfg, "model_use_robot_state", False):
                    r.states, r.state_t_stride = _lib.ptr(states), states.stride(0)
                r.masks, r.mask_t_stride = _lib.ptr(masks), masks.stride(0)
            if dev_robot is not None:
                c.robot = C.pointer(dev_robot.c_model)
                c.robot_start_state = _lib.ptr(start_norm)
                c.robot_render_masks = int(dev_robot.render_masks)
                c.robot_extra_radius = float(dev_robot.thick_extra)
            r.seed, r.noise_ctr_base = plan_seed, 0
            r.sample_mean = int(bool(getattr(cfg, "sample_mean", False)))
            r.zero_robot = int(_zero_robot(cfg))
            r.dontcare_cost = int(getattr(cfg, "reward_type", "") == "dontcare")
            r.sparse_cost = int(bool(getattr(cfg, "sparse_cost", False)))
            r.world_cost_weight = float(getattr(cfg, "world_cost_weight", 1.0))
            mean = torch.empty(L, 2, device=dev)
            std = torch.empty(L, 2, device=dev)
            elite = torch.empty(K, dtype=torch.int64, device=dev)
            costs = torch.empty(N, dtype=torch.float64, device=dev)
            _lib.check(self._lib.rac_cem_plan(m.handle, C.byref(c), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(elite),
                                              _lib.ptr(costs), _lib.stream_ptr()), m.handle, "rac_cem_plan")
            self.last_costs, self.last_elite_idx, self.last_std = costs, elite, std
            return mean

        # ---- per-iteration path: sharded candidates and / or a host-side robot model
        m.prepare(n_local)
        mean = torch.zeros(L, 2, device=dev)
        std = torch.ones(L, 2, device=dev) * float(self.init_std)
        act2 = torch.empty(N, L, 2, device=dev)
        act5 = torch.empty(n_local, L, A, device=dev)
        local_cost = torch.empty(n_local, dtype=torch.float64, device=dev)
        elite = torch.empty(K, dtype=torch.int64, device=dev)
        rollouts = None
        peer = parallel.PeerCostExchange.get(self, N, self.process_group, dev) if (world > 1 and not host_robot) else None
        for i in range(I):
            nz = noise[i] if noise is not None else None
            _lib.check(self._lib.rac_cem_sample(_lib.ptr(mean), _lib.ptr(std), _lib.ptr(nz), plan_seed, i, N, L, A, lo,
                                                n_local, 0.05, _lib.ptr(act2), _lib.ptr(act5), _lib.stream_ptr()),
                       None, "rac_cem_sample")
            last = i == I - 1
            if host_robot or (last and (opt_traj is not None or self.plot_rollouts)):
                ts.cand_offset = lo
                kw = {}
                if self.precomputed_robot is not None:
                    # this shard of the caller's device-resident states / masks (and the expert trajectory's, if given)
                    ps, pm = self.precomputed_robot
                    kw = dict(states=ps[:, lo:hi].contiguous(), masks=pm[:, lo:hi].contiguous(),
                              opt_robot=getattr(self, "precomputed_opt_robot", None))
                rollouts = ts.generate_model_rollouts(act5.cpu(), start, goal, opt_traj=opt_traj if last else None,
                                                      ret_obs=self.plot_rollouts and last, **kw)
                local_cost.copy_(torch.from_numpy(rollouts["sum_cost"]))
            else:
                states = masks = None
                if self.precomputed_robot is not None:
                    states, masks = self.precomputed_robot
                    # a shard of (T+1, N, ...) is contiguous per time step and strided across steps; the C ABI takes
                    # the time stride explicitly
                    states = _StridedView(states[:, lo:hi]) if getattr(cfg, "model_use_robot_state", False) else None
                    masks = _StridedView(masks[:, lo:hi])
                elif dev_robot is not None:
                    # robot_model.predict_batch on the device (trajectory_sampler.py:100-109), this shard only
                    states = dev_robot.predict_states(start_norm, act5)
                    masks = dev_robot.render(states, thick=True)
                    if not getattr(cfg, "model_use_robot_state", False):
                        states = None
                fused = peer is not None and not (last and (opt_traj is not None or self.plot_rollouts))
                try:
                    ts._rollout_device(act5, start_img, goal_imgs, goal_masks, states, masks, None, n_local, L, local_cost,
                                       cand_offset=lo, noise_ctr=i * L, peer=peer.target(lo) if fused else None)
                    if fused:
                        # the per-candidate costs are already in every rank's buffer (stored by the cost kernel through
                        # NVLink peer memory); one flag barrier orders them against the replicated top-k
                        costs = peer.finish(self._lib)
                except Exception:
                    # the exchange counts iterations on the host (buffer parity, flag sequence): a rank that failed
                    # between target() and finish() would meet the others one count off next time. Drop it; the next
                    # plan rendezvouses afresh (the other ranks of THIS plan are lost either way, as with any collective)
                    self._peer_exchange = None
                    raise
                if fused:
                    _lib.check(self._lib.rac_topk(_lib.ptr(costs), N, K, _lib.ptr(elite), None, _lib.stream_ptr()), None,
                               "rac_topk")
                    _lib.check(self._lib.rac_cem_refit(_lib.ptr(act2), L, _lib.ptr(elite), K, 0.001, _lib.ptr(mean),
                                                       _lib.ptr(std), _lib.stream_ptr()), None, "rac_cem_refit")
                    continue
            costs = parallel.all_gather_costs(local_cost, N, self.process_group)
            _lib.check(self._lib.rac_topk(_lib.ptr(costs), N, K, _lib.ptr(elite), None, _lib.stream_ptr()), None,
                       "rac_topk")
            _lib.check(self._lib.rac_cem_refit(_lib.ptr(act2), L, _lib.ptr(elite), K, 0.001, _lib.ptr(mean),
                                               _lib.ptr(std), _lib.stream_ptr()), None, "rac_cem_refit")
        self.last_costs, self.last_elite_idx, self.last_std = costs, elite, std
        self.last_rollouts = rollouts
        return mean


class _StridedView:
    """A (T+1, n_local, ...) slice of a (T+1, N, ...) tensor: contiguous per time step, strided across steps."""

    def __init__(self, t):
        self._t = t

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self._t.data_ptr()

    def stride(self, i):
        return self._t.stride(i)
