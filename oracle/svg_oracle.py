"""TEST INFRASTRUCTURE -- CPU oracle of the planning hot path. Not a product path: only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import this module.

A plain fp32 restatement (torch.nn.functional on CPU, no reference code) of what the reference computes on the path
  CEMPolicy.get_action -> TrajectorySampler.generate_model_rollouts -> SVGConvModel.forward -> RobotWorldCost
so that the CUDA kernels can be checked where /root/reference does not exist (the GPU box). The oracle is PINNED:
tests/test_oracle_golden.py compares every function below with outputs of the UNMODIFIED reference executed in the
build container (oracle/make_golden.py -> tests/golden/*.npz). The reference itself ships no tests or golden vectors
for this path, so reference-generated fixtures are the only pin available.

Each function cites the reference lines it restates (paths relative to the reference checkout).
"""
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

H_IMG, W_IMG = 48, 64


# --------------------------------------------------------------------------------------------- configuration
def make_cfg(g_dim=512, z_dim=64, action_dim=5, robot_dim=5, model_use_mask=False, model_use_future_mask=False,
             model_use_robot_state=False, model_use_future_robot_state=False, reconstruction_loss="l1",
             reward_type="weighted", last_frame_skip=True, sample_mean=False, sparse_cost=False,
             robot_cost_weight=0.0, world_cost_weight=1.0, black_robot_input=False, candidates_batch_size=200,
             topk=5, **extra):
    """The attributes of the reference argparse Namespace that the path reads (src/config/__init__.py:165-357)."""
    ns = SimpleNamespace(
        image_width=W_IMG, image_height=H_IMG, channels=3, g_dim=g_dim, z_dim=z_dim, action_dim=action_dim,
        robot_dim=robot_dim, model_use_mask=model_use_mask, model_use_future_mask=model_use_future_mask,
        model_use_robot_state=model_use_robot_state, model_use_future_robot_state=model_use_future_robot_state,
        model_use_heatmap=False, model_use_future_heatmap=False, reconstruction_loss=reconstruction_loss,
        reward_type=reward_type, last_frame_skip=last_frame_skip, sample_mean=sample_mean, sparse_cost=sparse_cost,
        robot_cost_weight=robot_cost_weight, world_cost_weight=world_cost_weight,
        black_robot_input=black_robot_input, candidates_batch_size=candidates_batch_size, topk=topk,
        lstm_group_norm=False, batch_size=16, device=torch.device("cpu"), debug_cem=False, img_cost_threshold=None,
        img_cost_world_norm=True,
    )
    for k, v in extra.items():
        setattr(ns, k, v)
    return ns


def encoder_in_channels(cfg):
    """dynamics.py:476-487."""
    c = cfg.channels
    if cfg.model_use_mask:
        c += 1
        if cfg.model_use_future_mask:
            c += 1
    return c


def state_dict_spec(cfg):
    """Key -> shape of SVGConvModel.state_dict() (dynamics.py:467-516, vgg_64.py:87-121,196-221, lstm.py:121-127,
    201-210,273-274). Order follows module registration order."""
    g, z, a, r = cfg.g_dim, cfg.z_dim, cfg.action_dim, cfg.robot_dim
    spec = OrderedDict()

    def vgg(prefix, cin, cout):
        spec[f"{prefix}.main.0.weight"] = (cout, cin, 3, 3)
        spec[f"{prefix}.main.1.weight"] = (cout,)
        spec[f"{prefix}.main.1.bias"] = (cout,)
        spec[f"{prefix}.main.1.running_mean"] = (cout,)
        spec[f"{prefix}.main.1.running_var"] = (cout,)
        spec[f"{prefix}.main.1.num_batches_tracked"] = ()

    def conv(prefix, cin, cout, k=3):
        spec[f"{prefix}.weight"] = (cout, cin, k, k)
        spec[f"{prefix}.bias"] = (cout,)

    def convlstm(prefix):
        for layer, k in ((0, 5), (1, 3)):
            if getattr(cfg, "lstm_group_norm", False):
                # NormConvLSTMCell (lstm.py:151-175): Sequential(Conv2d, GroupNorm(16, 4g)) x2 + GroupNorm(16, g)
                for gate in ("ih_gates", "hh_gates"):
                    conv(f"{prefix}.lstm.{layer}.{gate}.0", g, 4 * g, k)
                    spec[f"{prefix}.lstm.{layer}.{gate}.1.weight"] = (4 * g,)
                    spec[f"{prefix}.lstm.{layer}.{gate}.1.bias"] = (4 * g,)
                spec[f"{prefix}.lstm.{layer}.c_norm.weight"] = (g,)
                spec[f"{prefix}.lstm.{layer}.c_norm.bias"] = (g,)
            else:
                conv(f"{prefix}.lstm.{layer}.gates", 2 * g, 4 * g, k)

    nc = encoder_in_channels(cfg)
    for name, cin, cout in [("c1.0", nc, 64), ("c1.1", 64, 64), ("c2.0", 64, 128), ("c2.1", 128, 128),
                            ("c3.0", 128, 256), ("c3.1", 256, 256), ("c3.2", 256, 256), ("c4.0", 256, 512),
                            ("c4.1", 512, 512), ("c4.2", 512, g)]:
        vgg(f"encoder.{name}", cin, cout)
    lstm_c = g + a + z
    post_c = g
    prior_c = g + a
    if cfg.model_use_robot_state:
        lstm_c += r
        post_c += r
        prior_c += r
    if cfg.model_use_future_robot_state:
        lstm_c += r
        prior_c += r
    conv("frame_pred_input_conv", lstm_c, g)
    convlstm("frame_predictor")
    conv("posterior_input_conv", post_c, g)
    conv("prior_input_conv", prior_c, g)
    for p in ("posterior", "prior"):
        convlstm(p)
        conv(f"{p}.mu_net", g, z)
        conv(f"{p}.logvar_net", g, z)
    for name, cin, cout in [("upc2.0", g, 512), ("upc2.1", 512, 512), ("upc2.2", 512, 256), ("upc3.0", 512, 256),
                            ("upc3.1", 256, 256), ("upc3.2", 256, 128), ("upc4.0", 256, 128), ("upc4.1", 128, 64),
                            ("upc5.0", 128, 64)]:
        vgg(f"decoder.{name}", cin, cout)
    spec["decoder.upc5.1.weight"] = (64, cfg.channels + 1, 3, 3)  # ConvTranspose2d weight is (in, out, k, k)
    spec["decoder.upc5.1.bias"] = (cfg.channels + 1,)
    return spec


def make_state_dict(cfg, seed=0):
    """Deterministic synthetic weights with the reference's key set. Conv weights ~ N(0, 0.02) as the reference's
    init_weights (base.py:26-36); biases, BN affine and BN running statistics are randomised so that folding / bias
    bugs are visible. CPU generator => bit-identical on every box with the same torch."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(int(seed))
    sd = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=gen) + 0.5
        elif key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=gen) * 0.1
        elif ".main.1.weight" in key or "_gates.1.weight" in key or key.endswith("c_norm.weight"):
            sd[key] = torch.rand(shape, generator=gen) + 0.5
        elif ".main.1.bias" in key:
            sd[key] = torch.randn(shape, generator=gen) * 0.1
        elif key.endswith(".bias"):
            sd[key] = torch.randn(shape, generator=gen) * 0.05
        else:
            sd[key] = torch.randn(shape, generator=gen) * 0.02
    return sd


def trained_like(sd, gate_scale0=6.0, gate_scale1=12.0, head_scale=8.0):
    """A "trained-like" variant of a synthetic state dict (SURVEY.md 8(d) caveat: random init is the forgiving case --
    gate pre-activations of std 0.5-1 and a decoder head that outputs a nearly flat 0.5 +- 0.02 image). The ConvLSTM gate
    weights are scaled so that the pre-activations have magnitude ~3-5 (saturating sigmoids / tanh, sharp gates) and
    the final ConvTranspose is scaled so that predicted pixels and the compositing mask span a real range. A
    deterministic transform of make_state_dict's output: the reference and the CUDA path both load the result."""
    out = OrderedDict()
    for k, v in sd.items():
        if ".lstm.0.gates." in k:
            v = v * gate_scale0
        elif ".lstm.1.gates." in k:
            v = v * gate_scale1
        elif k == "decoder.upc5.1.weight":
            v = v * head_scale
        out[k] = v
    return out


# --------------------------------------------------------------------------------------------- model
class SVGOracle:
    """Functional fp32 restatement of SVGConvModel in eval mode (dynamics.py:457-644)."""

    def __init__(self, cfg, state_dict):
        self.cfg = cfg
        self.sd = {k: v.detach().to(torch.float32) if v.is_floating_point() else v for k, v in state_dict.items()}
        self.hidden = None
        self.trace = None  # set to {} to record intermediate activations (NCHW) for layer-by-layer GPU diagnosis
        self.bn_training = False  # True: BatchNorm uses batch statistics and updates running stats (trainer.py:754)
        # emulate_bf16: round conv weights and inter-layer activations to bf16 (straight-through for autograd) at the
        # points where the CUDA path stores bf16, so that ReLU / max-pool / |.| decisions agree with it. Used only to
        # separate rounding-induced gradient noise from implementation errors in the training tests.
        self.emulate_bf16 = False

    def _q(self, x):
        if not self.emulate_bf16:
            return x
        return x + (x.to(torch.bfloat16).float() - x).detach()

    def _rec(self, name, t):
        if self.trace is not None:
            if t.requires_grad:
                t.retain_grad()  # training diagnosis: intermediate gradients are compared layer by layer
                self.trace[name] = t
            else:
                self.trace[name] = t.clone()
        return t

    # vgg_layer: conv3x3(no bias) -> BatchNorm2d(eval) -> LeakyReLU(0.2)  (vgg_64.py:8-18)
    def _vgg(self, x, prefix):
        sd = self.sd
        x = F.conv2d(x, self._q(sd[f"{prefix}.main.0.weight"]), None, 1, 1)
        x = F.batch_norm(x, sd[f"{prefix}.main.1.running_mean"], sd[f"{prefix}.main.1.running_var"],
                         sd[f"{prefix}.main.1.weight"], sd[f"{prefix}.main.1.bias"], self.bn_training, 0.1, 1e-5)
        return self._q(F.leaky_relu(x, 0.2))

    # ConvEncoder.forward (vgg_64.py:122-129)
    def encode(self, x):
        h1 = self._vgg(self._rec("a1", self._vgg(x, "encoder.c1.0")), "encoder.c1.1")
        h2 = self._vgg(self._rec("a2", self._vgg(F.max_pool2d(h1, 2, 2), "encoder.c2.0")), "encoder.c2.1")
        h3 = F.max_pool2d(h2, 2, 2)
        for i in range(3):
            h3 = self._vgg(h3, f"encoder.c3.{i}")
        h4 = F.max_pool2d(h3, 2, 2)
        for i in range(3):
            h4 = self._vgg(h4, f"encoder.c4.{i}")
        for n_, t_ in (("h1", h1), ("h2", h2), ("h3", h3), ("h4", h4)):
            self._rec(n_, t_)
        return h4, [h1, h2, h3, h4]

    # ConvDecoder.forward (vgg_64.py:223-241)
    def decode(self, vec, skip):
        sd = self.sd
        d = vec
        for i in range(3):
            d = self._vgg(d, f"decoder.upc2.{i}")
            self._rec(f"d2.{i}", d)
        d = self._rec("cat3", torch.cat([F.interpolate(d, scale_factor=2, mode="nearest"), skip[2]], 1))
        for i in range(3):
            d = self._rec(f"d3.{i}", self._vgg(d, f"decoder.upc3.{i}"))
        d = self._rec("cat4", torch.cat([F.interpolate(d, scale_factor=2, mode="nearest"), skip[1]], 1))
        for i in range(2):
            d = self._rec(f"d4.{i}", self._vgg(d, f"decoder.upc4.{i}"))
        d = self._rec("cat5", torch.cat([F.interpolate(d, scale_factor=2, mode="nearest"), skip[0]], 1))
        d = self._rec("d5", self._vgg(d, "decoder.upc5.0"))
        d = F.conv_transpose2d(d, self._q(sd["decoder.upc5.1.weight"]), sd["decoder.upc5.1.bias"], 1, 1)
        return torch.sigmoid(d)

    # ConvLSTM.init_hidden (lstm.py:218-250), SVGConvModel.init_hidden (dynamics.py:536-542)
    def init_hidden(self, batch_size):
        g = self.cfg.g_dim
        zeros = lambda: torch.zeros(batch_size, g, H_IMG // 8, W_IMG // 8)
        self.hidden = {name: [(zeros(), zeros()), (zeros(), zeros())]
                       for name in ("frame_predictor", "posterior", "prior")}

    # ConvLSTMCell.forward (lstm.py:129-149) x2, ConvLSTM.forward (lstm.py:252-257)
    def _convlstm(self, x, name):
        sd = self.sd
        norm = getattr(self.cfg, "lstm_group_norm", False)
        for layer, pad in ((0, 2), (1, 1)):
            h_prev, c_prev = self.hidden[name][layer]
            p = f"{name}.lstm.{layer}"
            if norm:
                # NormConvLSTMCell.forward (lstm.py:177-198): GroupNorm(16) on each gate convolution and on the cell
                gates = sum(F.group_norm(F.conv2d(t, self._q(sd[f"{p}.{gk}.0.weight"]), sd[f"{p}.{gk}.0.bias"], 1, pad),
                                         16, sd[f"{p}.{gk}.1.weight"], sd[f"{p}.{gk}.1.bias"], 1e-5)
                            for gk, t in (("ih_gates", x), ("hh_gates", h_prev)))
            else:
                gates = F.conv2d(torch.cat([x, h_prev], 1), self._q(sd[f"{p}.gates.weight"]), sd[f"{p}.gates.bias"], 1, pad)
            i, f, o, g_ = gates.chunk(4, 1)
            c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g_)
            if norm:
                c = F.group_norm(c, 16, sd[f"{p}.c_norm.weight"], sd[f"{p}.c_norm.bias"], 1e-5)
            h = self._q(torch.sigmoid(o) * torch.tanh(c))
            self.hidden[name][layer] = (h, c)
            self._rec(f"{name}.h{layer}", h)
            self._rec(f"{name}.c{layer}", c)
            x = h
        return x

    # GaussianConvLSTM.forward / reparameterize (lstm.py:276-286); eps is injected instead of drawn
    def _gaussian(self, x, name, eps):
        sd = self.sd
        h = self._convlstm(x, name)
        mu = F.conv2d(h, self._q(sd[f"{name}.mu_net.weight"]), sd[f"{name}.mu_net.bias"], 1, 1)
        logvar = F.conv2d(h, self._q(sd[f"{name}.logvar_net.weight"]), sd[f"{name}.logvar_net.bias"], 1, 1)
        z = self._q(eps * torch.exp(0.5 * logvar) + mu)
        return z, mu, logvar

    @staticmethod
    def _tile(v):
        return v[:, :, None, None].expand(-1, -1, H_IMG // 8, W_IMG // 8)

    @torch.no_grad()
    def forward(self, *args, **kw):
        """SVGConvModel.forward (dynamics.py:544-644) without autograd; see forward_grad for the arguments."""
        return self.forward_grad(*args, **kw)

    def forward_grad(self, image, mask, robot, action, eps, next_robot=None, eps_post=None, use_posterior=False,
                     force_use_prior=False, sample_mean=False, skip=None):
        """SVGConvModel.forward (dynamics.py:544-644). `robot` is a tensor or an (r, r_next) tuple (:596-597).
        Returns (x_pred, skip, mu, logvar, mu_p, logvar_p)."""
        cfg, sd = self.cfg, self.sd
        img = torch.cat([image, mask], 1) if cfg.model_use_mask else image
        h, curr_skip = self.encode(img)
        if cfg.last_frame_skip or skip is None:
            skip = curr_skip
        parts = [self._q(self._tile(action))]
        if cfg.model_use_robot_state:
            if cfg.model_use_future_robot_state:
                parts += [self._q(self._tile(robot[0])), self._q(self._tile(robot[1]))]
            else:
                parts += [self._q(self._tile(robot))]
        prior_in = self._q(F.conv2d(torch.cat(parts + [h], 1), self._q(sd["prior_input_conv.weight"]),
                                    sd["prior_input_conv.bias"], 1, 1))
        self._rec("prior_in", prior_in)
        z_p, mu_p, logvar_p = self._gaussian(prior_in, "prior", eps)
        z = mu_p if sample_mean else z_p
        mu = logvar = None
        if use_posterior:
            # dynamics.py:619 encodes `img` (the CURRENT frame) again: same values as h; in train mode the second
            # pass updates the BatchNorm running statistics a second time and is a second autograd path
            h_t = self.encode(img)[0] if self.bn_training else h
            post_parts = [self._q(self._tile(next_robot))] if cfg.model_use_robot_state else []
            post_in = self._q(F.conv2d(torch.cat(post_parts + [h_t], 1), self._q(sd["posterior_input_conv.weight"]),
                                       sd["posterior_input_conv.bias"], 1, 1))
            z_t, mu, logvar = self._gaussian(post_in, "posterior", eps_post)
            if not force_use_prior:
                z = z_t
        frame_in = self._q(F.conv2d(torch.cat(parts + [h, z], 1), self._q(sd["frame_pred_input_conv.weight"]),
                                    sd["frame_pred_input_conv.bias"], 1, 1))
        self._rec("z", z)
        self._rec("frame_in", frame_in)
        h_pred = self._convlstm(frame_in, "frame_predictor")
        x_pred = self.decode(h_pred, skip)
        return x_pred, skip, mu, logvar, mu_p, logvar_p


# --------------------------------------------------------------------------------------------- costs / criteria
def zero_robot_region(mask, image):
    """src/utils/image.py:5-20."""
    return torch.where(mask.bool().expand(-1, 3, -1, -1), torch.zeros_like(image), image)


def img_l2_cost(curr, goal):
    """ImgL2Cost._call_tensor (losses.py:224-235): -sqrt(sum((255 (curr-goal))^2)) per candidate."""
    d = (255 * (curr - goal)) ** 2
    return -(d.sum((1, 2, 3)).sqrt()).numpy()


def img_dontcare_cost(curr, goal, curr_mask, goal_mask):
    """ImgDontcareCost._call_tensor (losses.py:244-263): robot pixels of either mask ignored; divided by the number
    of world PIXELS (not x3, no +1)."""
    m2 = curr_mask.bool() | goal_mask.bool()
    d = (255 * (curr - goal)) ** 2
    d = torch.where(m2.expand(-1, 3, -1, -1), torch.zeros_like(d), d)
    dist = d.sum((1, 2, 3)).sqrt() / (~m2).sum((1, 2, 3))
    return -dist.numpy()


def l1_criterion(pred, target, batch_weight=None):
    """losses.py:13-19."""
    if batch_weight is not None:
        return (batch_weight * (target - pred).abs().mean((1, 2, 3))).mean()
    return (target - pred).abs().mean()


def dontcare_l1_criterion(pred, target, mask, robot_weight, batch_weight=None):
    """losses.py:35-50 (3-channel world-pixel count, +1)."""
    m3 = mask.bool().expand(-1, 3, -1, -1)
    diff = target - pred
    diff = torch.where(m3, diff * robot_weight, diff)
    world = (~m3).sum((1, 2, 3)) + 1
    if batch_weight is not None:
        return (batch_weight * diff.abs().sum((1, 2, 3)) / world).mean()
    return (diff.abs().sum((1, 2, 3)) / world).mean()


def mse_criterion(pred, target):
    """nn.MSELoss() (losses.py:11)."""
    return ((pred - target) ** 2).mean()


def dontcare_mse_criterion(pred, target, mask, robot_weight):
    """losses.py:21-33."""
    m3 = mask.bool().expand(-1, 3, -1, -1)
    diff = target - pred
    diff = torch.where(m3, diff * robot_weight, diff)
    world = (~m3).sum((1, 2, 3)) + 1
    return ((diff ** 2).sum((1, 2, 3)) / world).mean()


def robot_world_mse(pred, target, mask):
    """robot_mse_criterion / world_mse_criterion (losses.py:52-78)."""
    m3 = mask.bool().expand(-1, 3, -1, -1)
    d2 = (target - pred) ** 2
    robot = (torch.where(m3, d2, torch.zeros_like(d2)).sum((1, 2, 3)) / (m3.sum((1, 2, 3)) + 1)).mean()
    world = (torch.where(m3, torch.zeros_like(d2), d2).sum((1, 2, 3)) / ((~m3).sum((1, 2, 3)) + 1)).mean()
    return robot, world


def kl_criterion(mu1, logvar1, mu2, logvar2, bs):
    """losses.py:97-106."""
    s1, s2 = torch.exp(0.5 * logvar1), torch.exp(0.5 * logvar2)
    kld = torch.log(s2 / s1) + (torch.exp(logvar1) + (mu1 - mu2) ** 2) / (2 * torch.exp(logvar2)) - 0.5
    return kld.sum() / bs


# --------------------------------------------------------------------------------------------- rollout
@torch.no_grad()
def rollout_cost(model, cfg, actions, start_img_u8, goal_imgs_u8, goal_masks=None, states=None, masks=None,
                 eps=None, ret_obs=False):
    """TrajectorySampler.generate_model_rollouts (trajectory_sampler.py:35-199) for given robot states / masks
    (the MuJoCo `predict_batch` is an input, :100-109). actions (N, L, A); eps (L, N, z, 6, 8).
    Returns dict(sum_cost float64[N], step_cost float32[L, N], obs (N, L, 3, H, W) if ret_obs)."""
    N, L = actions.shape[0], actions.shape[1]
    sum_cost = np.zeros(N)
    step_cost = np.zeros((L, N), dtype=np.float32)
    obs = torch.zeros(N, L, 3, H_IMG, W_IMG) if ret_obs else None
    goal_imgs = torch.stack([torch.from_numpy(np.ascontiguousarray(g)).permute(2, 0, 1).float() / 255
                             for g in goal_imgs_u8])
    gmasks = torch.stack([torch.from_numpy(np.ascontiguousarray(g)) for g in goal_masks]) if goal_masks is not None else None
    zero_robot = ("dontcare" in cfg.reconstruction_loss) or cfg.black_robot_input
    use_mask_cost = zero_robot or ("dontcare" in cfg.reward_type)
    model.init_hidden(N)
    curr = torch.from_numpy(start_img_u8.copy()).permute(2, 0, 1).float() / 255
    curr = curr.expand(N, -1, -1, -1)
    for t in range(L):
        mask = masks[t] if cfg.model_use_mask else None
        state = states[t] if cfg.model_use_robot_state else None
        if zero_robot:
            curr = zero_robot_region(masks[t], curr)
        if cfg.model_use_future_mask:
            mask = torch.cat([mask, masks[t + 1]], 1)
        if cfg.model_use_future_robot_state:
            state = (state, states[t + 1])
        x_pred = model.forward(curr, mask, state, actions[:, t], eps[t], sample_mean=cfg.sample_mean)[0]
        rgb, m = x_pred[:, :3], x_pred[:, 3:4]
        nxt = (1 - m) * curr + m * rgb
        if zero_robot:
            nxt = zero_robot_region(masks[t + 1], nxt)
        gi = t if t < len(goal_imgs) else -1
        rew = 0.0
        if (not cfg.sparse_cost) or t == L - 1:
            if cfg.world_cost_weight != 0:
                if cfg.reward_type == "dontcare":
                    c = img_dontcare_cost(nxt, goal_imgs[gi], masks[t + 1], gmasks[gi])
                else:
                    c = img_l2_cost(nxt, goal_imgs[gi])
                rew = cfg.world_cost_weight * c
        sum_cost += rew
        step_cost[t] = rew
        if ret_obs:
            obs[:, t] = nxt
        curr = nxt
    out = {"sum_cost": sum_cost, "step_cost": step_cost}
    if ret_obs:
        out["obs"] = obs.numpy()
    return out


# --------------------------------------------------------------------------------------------- CEM
def topk_largest(costs, k):
    """costs.topk(K) (cem.py:97) with the tie rule of SURVEY.md 8(a) A9: value descending, lowest index first."""
    c = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(c)), -c))
    return order[:k].astype(np.int64)


def cem_sample(mean, std, noise, it, clamp=0.05):
    """cem.py:80-85: Normal(mean, std).sample == mean + std * n(0,1); last candidate zeroed at iteration 0."""
    act = mean[None] + std[None] * noise
    if it == 0:
        act[-1] = 0
    return act.clamp(-clamp, clamp)


def cem_refit(act_seq, elite_idx, std_floor=0.001):
    """cem.py:98-104: unbiased std / mean over the elites, std floored."""
    top = act_seq[torch.as_tensor(elite_idx)]
    std, mean = torch.std_mean(top, dim=0)
    return mean, torch.max(std_floor * torch.ones_like(std), std)


@torch.no_grad()
def cem_plan(model, cfg, noise, topk, init_std, start_img_u8, goal_imgs_u8, goal_masks=None, robot_fn=None, eps=None,
             action_dim=None):
    """CEMPolicy.get_action (cem.py:56-111) with injected sampling noise (I, N, L, 2) and z noise (I, L, N, z, 6, 8).
    robot_fn(actions5) -> (states, masks) stands in for robot_model.predict_batch. Returns (mean, history)."""
    I, N, L, _ = noise.shape
    A = action_dim or cfg.action_dim
    mean = torch.zeros(L, 2)
    std = torch.ones(L, 2) * init_std
    hist = []
    for it in range(I):
        act = cem_sample(mean, std, noise[it].clone(), it)
        padded = torch.cat([act, torch.zeros(N, L, A - 2)], 2)
        states, masks = robot_fn(padded) if robot_fn is not None else (None, None)
        r = rollout_cost(model, cfg, padded, start_img_u8, goal_imgs_u8, goal_masks, states, masks,
                         eps[it] if eps is not None else None)
        idx = topk_largest(r["sum_cost"], topk)
        mean, std = cem_refit(act, idx)
        hist.append({"sum_cost": r["sum_cost"].copy(), "elite": idx.copy(), "mean": mean.numpy().copy(),
                     "std": std.numpy().copy(), "act": act.numpy().copy()})
    return mean.numpy(), hist
