"""Planning costs and training criteria with the reference's names and call signatures
(src/prediction/losses.py:13-50,97-106,181-335), computed by the CUDA kernels in csrc/misc_kernels.cu.

The batched tensor path (what the planner uses) runs on the GPU; like the reference it returns a float32 numpy array
of NEGATIVE distances (one device->host read, losses.py:234,262). Inside `TrajectorySampler` the same cost is fused
into the decoder epilogue and never leaves the device (csrc/epilogue.cuh::epi_frame)."""
import numpy as np
import torch

from . import _lib
from .state import State


def _dev(t):
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def _masked_cost(curr_img, goal_img, curr_mask, goal_mask, dontcare):
    if curr_img.dim() == 3:  # single image version (losses.py:229-230)
        out = _masked_cost(curr_img[None], goal_img, None if curr_mask is None else curr_mask[None], goal_mask, dontcare)
        return out[0]
    if curr_img.dim() != 4:
        raise NotImplementedError(f"Tensor shape {tuple(curr_img.shape)} not supported")
    lib = _lib.load()
    n, ch, h, w = curr_img.shape
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    curr = _dev(curr_img)
    goal = _dev(goal_img.expand(ch, h, w) if goal_img.dim() == 3 else goal_img)
    cm = _dev(curr_mask) if dontcare else None
    gm = _dev(goal_mask) if dontcare else None
    out = torch.empty(n, device="cuda", dtype=torch.float32)
    _lib.check(lib.rac_masked_cost(_lib.ptr(curr), _lib.ptr(goal), _lib.ptr(cm), _lib.ptr(gm), int(dontcare),
                                   _lib.ptr(out), n, h * w, _lib.stream_ptr()), None, "rac_masked_cost")
    return out.cpu().numpy()


class Cost:
    def __init__(self, config):
        self._config = config

    def __call__(self, curr: State, goal: State):
        raise NotImplementedError()


class RobotL2Cost(Cost):
    """losses.py:181-206. In CEM `State.state` is None, so this contributes 0.0 (losses.py:189-190)."""
    name = "robot_l2"

    def __call__(self, curr: State, goal: State):
        if curr.state is None or goal.state is None:
            return 0.0
        if isinstance(curr.state, torch.Tensor) or isinstance(goal.state, torch.Tensor):
            d = (torch.as_tensor(curr.state) - torch.as_tensor(goal.state)) ** 2
            if d.dim() == 2:
                s = d.sum(1)
            elif d.dim() == 1:
                s = d.sum()
            else:
                raise NotImplementedError(f"Tensor shape {tuple(d.shape)} not supported")
            return -s.sqrt().cpu().numpy()
        return -np.linalg.norm(np.asarray(curr.state) - np.asarray(goal.state))


class ImgL2Cost(Cost):
    """losses.py:209-240."""
    name = "img_l2"

    def __call__(self, curr: State, goal: State):
        if curr.img is None or goal.img is None:
            return 0
        if not (isinstance(curr.img, torch.Tensor) or isinstance(goal.img, torch.Tensor)):
            raise NotImplementedError("numpy image costs belong to the simulator path (out of scope, SURVEY.md 2.1)")
        return _masked_cost(torch.as_tensor(curr.img), torch.as_tensor(goal.img), None, None, False)


class ImgDontcareCost(Cost):
    """losses.py:242-288."""
    name = "img_dontcare"

    def __call__(self, curr: State, goal: State):
        if curr.img is None or goal.img is None:
            return 0
        if not (isinstance(curr.img, torch.Tensor) or isinstance(goal.img, torch.Tensor)):
            raise NotImplementedError("numpy image costs belong to the simulator path (out of scope, SURVEY.md 2.1)")
        return _masked_cost(torch.as_tensor(curr.img), torch.as_tensor(goal.img), curr.mask, goal.mask, True)


class RobotWorldCost(Cost):
    """Combination of a robot and a world cost (losses.py:290-335)."""

    def __init__(self, config):
        self._config = config
        self.robot_cost_weight = getattr(config, "robot_cost_weight", 0.0)
        self.robot_cost = RobotL2Cost(config)
        self.world_cost_weight = getattr(config, "world_cost_weight", 1.0)
        self.world_cost = ImgL2Cost(config)
        if getattr(config, "reward_type", "weighted") == "dontcare":
            self.world_cost = ImgDontcareCost(config)

    def __call__(self, curr: State, goal: State, print_cost=False, return_info=False):
        total, info, text = 0, {}, ""
        for w, c in ((self.robot_cost_weight, self.robot_cost), (self.world_cost_weight, self.world_cost)):
            if w == 0:
                continue
            cost = w * c(curr, goal)
            if return_info:
                if type(cost) in (np.float64, float):
                    info[c.name] = cost
                else:
                    raise NotImplementedError()
            if print_cost:
                vals = [cost] if type(cost) in (np.float64, float) else list(cost)
                text += "".join(f" {c.name}: {v:.4f} ," for v in vals)
            total += cost
        if print_cost:
            print(text)
        return (total, info) if return_info else total


# ---- training criteria: forward values (losses.py:13-19,35-50,97-106) ----
def _recon(kind, prediction, target, mask=None, robot_weight=0.0, batch_weight=None):
    """rac_recon_loss: the four criteria of PredictionTrainer._recon_loss (trainer.py:149-161) on (B,3,H,W) tensors."""
    lib = _lib.load()
    p, t = _dev(prediction), _dev(target)
    m = _dev(mask) if mask is not None else None
    bw = _dev(batch_weight).reshape(-1) if batch_weight is not None else None
    n, _, h, w = p.shape
    if bw is not None and bw.numel() != n:
        raise ValueError(f"batch_weight has {bw.numel()} entries for a batch of {n}")
    scratch = torch.empty(n + 1, device="cuda")
    _lib.check(lib.rac_recon_loss(_lib.ptr(p), _lib.ptr(t), _lib.ptr(m), _lib.ptr(bw), kind, float(robot_weight),
                                  _lib.ptr(scratch), C_void(scratch, n), n, h * w, _lib.stream_ptr()), None,
               "rac_recon_loss")
    return scratch[n]


def C_void(t, index):
    import ctypes

    return ctypes.c_void_p(t.data_ptr() + index * t.element_size())


def mse_criterion(prediction, target):
    """nn.MSELoss() (losses.py:11), the reference's default reconstruction loss (src/config/__init__.py:235-238)."""
    return _recon(2, prediction, target)


def dontcare_mse_criterion(prediction, target, mask, robot_weight):
    """losses.py:21-33."""
    return _recon(3, prediction, target, mask, robot_weight)


def l1_criterion(prediction, target, batch_weight=None):
    if batch_weight is not None:  # movement weighting (losses.py:17-18)
        return _recon(0, prediction, target, batch_weight=batch_weight)
    lib = _lib.load()
    p, t = _dev(prediction), _dev(target)
    out = torch.empty(1, device="cuda")
    _lib.check(lib.rac_l1_loss(_lib.ptr(p), _lib.ptr(t), _lib.ptr(out), p.numel(), _lib.stream_ptr()), None, "rac_l1_loss")
    return out[0]


def dontcare_l1_criterion(prediction, target, mask, robot_weight, batch_weight=None):
    if batch_weight is not None:  # losses.py:48-49
        return _recon(1, prediction, target, mask, robot_weight, batch_weight)
    lib = _lib.load()
    p, t, m = _dev(prediction), _dev(target), _dev(mask)
    n, _, h, w = p.shape
    out = torch.empty(1, device="cuda")
    _lib.check(lib.rac_dontcare_l1_loss(_lib.ptr(p), _lib.ptr(t), _lib.ptr(m), float(robot_weight), _lib.ptr(out), n,
                                        h * w, _lib.stream_ptr()), None, "rac_dontcare_l1_loss")
    return out[0]


def _robot_world_mse(prediction, target, mask):
    lib = _lib.load()
    p, t, m = _dev(prediction), _dev(target), _dev(mask)
    n, _, h, w = p.shape
    out = torch.zeros(2, device="cuda")
    _lib.check(lib.rac_robot_world_mse(_lib.ptr(p), _lib.ptr(t), _lib.ptr(m), _lib.ptr(out), n, h * w,
                                       _lib.stream_ptr()), None, "rac_robot_world_mse")
    return out


def robot_mse_criterion(prediction, target, mask):
    """losses.py:52-64."""
    return _robot_world_mse(prediction, target, mask)[0]


def world_mse_criterion(prediction, target, mask):
    """losses.py:66-78."""
    return _robot_world_mse(prediction, target, mask)[1]


def kl_criterion(mu1, logvar1, mu2, logvar2, bs):
    assert mu1.shape[0] == bs, f"{mu1.shape[0]} != {bs}"
    lib = _lib.load()
    a, b, c, d = _dev(mu1), _dev(logvar1), _dev(mu2), _dev(logvar2)
    out = torch.empty(1, device="cuda")
    _lib.check(lib.rac_kl_loss(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(d), _lib.ptr(out), a.numel(), int(bs),
                               _lib.stream_ptr()), None, "rac_kl_loss")
    return out[0]
