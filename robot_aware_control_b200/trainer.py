"""SVG training step on B200: the arithmetic of the reference `PredictionTrainer._train_step`
(src/prediction/trainer.py:326-465) + Adam (`_init_models`, :109-122) through libracb200.so.

    trainer = SVGTrainer(cfg, model)                  # model: robot_aware_control_b200.SVGConvModel
    losses = trainer.train_step(batch)                # batch as the reference loader yields it, time-first

`batch` = {"images": (T,B,3,H,W), "masks": (T,B,1,H,W), "states": (T,B,5), "actions": (T-1,B,A)} (reference
robonet_dataset.py:434-451). Returned dict = the reference's logged losses (divided by n_future, trainer.py:463-464).

Parameters, BatchNorm running statistics, gradients and Adam moments live in flat fp32 CUDA tensors; the model's
nn.Parameters are VIEWS into them, so `model.state_dict()` / checkpoints always see the trained values.
Data parallel (new, the reference is single-process): pass `process_group`; every rank runs its own batch of B, the
flat gradient buffer is all-reduced (NCCL, SUM; the 1 / world factor is applied inside the Adam kernel) -- the ConvLSTM
gate convolutions (89 % of the parameters, whose BPTT finishes long before the encoder's) as soon as their gradient is
complete, underneath the rest of the backward pass, the remainder afterwards. BatchNorm statistics stay per rank.

Scheduled sampling follows the reference (same probability schedule, same global numpy generator); when the model's
own prediction is fed back, its gradient flows into the previous step as in the reference (`x_pred.clone()`).
`cfg.lstm_group_norm` (NormConvLSTMCell, the cell of the authors' deployed checkpoints) trains as well.
All four `cfg.reconstruction_loss` kinds (l1, dontcare_l1, mse -- the argparse default --, dontcare_mse) and the
movement weighting (`cfg.load_movement_info`, batch["high_movement"]) are implemented. Limits (raise): heatmaps,
multiview; the batch must be a multiple of 4 clips.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, pack
from .config import svg_config_from


def _layer_tables(model, offsets, boffsets):
    """For each packed layer: where its tensors live in the flat buffers + the packed-column / packed-channel maps of
    pack.py expressed as offsets (see rac_train_layer in include/racb200.h)."""
    c = model._c
    g, z, a, r = c.g_dim, c.z_dim, c.action_dim, c.robot_dim
    use_r = bool(c.model_use_robot_state)
    use_r2 = use_r and bool(c.model_use_future_robot_state)
    naux = a + (r if use_r else 0) + (r if use_r2 else 0)
    sd_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    out = {}

    def cols(splits, k2, in_stride=None):
        res, ci = [], 0
        for real, padded in splits:
            res += [(ci + i) * (in_stride or k2) for i in range(real)] + [-1] * (padded - real)
            ci += real
        return res

    def conv_entry(prefix, splits, col_src, n_packed, bias=True, bn=None):
        w_key = f"{prefix}.weight" if bn is None else f"{prefix}.main.0.weight"
        cout, cin, k, _ = sd_shapes[w_key]
        k2 = k * k
        w_off = offsets[w_key]
        rows = [(w_off + cs * cin * k2) if cs >= 0 else -1 for cs in col_src] + [-1] * (n_packed - len(col_src))
        b = None
        if bias:
            b_off = offsets[f"{prefix}.bias"]
            b = [(b_off + cs) if cs >= 0 else -1 for cs in col_src] + [-1] * (n_packed - len(col_src))
        e = dict(row_off=rows, col_off=cols(splits, k2), bias_off=b, flip=0, w_off=w_off)
        # the weight (and, when it directly follows in the flat buffer, the bias) of this convolution: nothing else
        # writes that gradient range, so it is final once the layer's weight gradient has been unpacked
        cnt = cout * cin * k2
        if bias and offsets[f"{prefix}.bias"] == w_off + cnt:
            cnt += cout
        e.update(grad_off=w_off, grad_count=cnt, w_count=cout * cin * k2)
        if bn is not None:
            e.update(gamma_off=offsets[f"{prefix}.main.1.weight"], beta_off=offsets[f"{prefix}.main.1.bias"],
                     rmean_off=boffsets[f"{prefix}.main.1.running_mean"], rvar_off=boffsets[f"{prefix}.main.1.running_var"])
        return e

    rup = lambda x, m: (x + m - 1) // m * m
    out["ENC_C1_0"] = conv_entry("encoder.c1.0", [(sd_shapes["encoder.c1.0.main.0.weight"][1],) * 2], list(range(64)), 64,
                                 bias=False, bn=True)
    for name, prefix in pack._VGG.items():
        cout, cin = sd_shapes[f"{prefix}.main.0.weight"][:2]
        out[name] = conv_entry(prefix, [(cin, cin)], list(range(cout)), rup(cout, 64 if cout == 64 else 128), bias=False, bn=True)
    out["PRIOR_IN"] = conv_entry("prior_input_conv", [(naux, 64), (g, g)], list(range(g)), rup(g, 128))
    out["FP_IN"] = conv_entry("frame_pred_input_conv", [(naux, 64), (g, g), (z, 64)], list(range(g)), rup(g, 128))
    out["POST_IN"] = conv_entry("posterior_input_conv", ([(r, 64)] if use_r else []) + [(g, g)], list(range(g)), rup(g, 128))
    gate_cols = [gate * g + ch for ch in range(g) for gate in range(4)]
    for tag, prefix in (("PRIOR", "prior"), ("POST", "posterior"), ("FP", "frame_predictor")):
        for layer in (0, 1):
            p = f"{prefix}.lstm.{layer}"
            if not c.lstm_group_norm:
                out[f"{tag}_LSTM{layer}"] = conv_entry(f"{p}.gates", [(g, g), (g, g)], gate_cols, rup(4 * g, 128))
                continue
            # NormConvLSTMCell (lstm.py:151-198): separate ih / hh convolutions, each followed by GroupNorm(16, 4g)
            # (its affine offsets ride in gamma_off / beta_off), and the cell's c_norm on the ih entry
            ih = conv_entry(f"{p}.ih_gates.0", [(g, g)], gate_cols, rup(4 * g, 128))
            ih.update(gamma_off=offsets[f"{p}.ih_gates.1.weight"], beta_off=offsets[f"{p}.ih_gates.1.bias"],
                      cnorm_gamma_off=offsets[f"{p}.c_norm.weight"], cnorm_beta_off=offsets[f"{p}.c_norm.bias"])
            hh = conv_entry(f"{p}.hh_gates.0", [(g, g)], gate_cols, rup(4 * g, 128))
            hh.update(gamma_off=offsets[f"{p}.hh_gates.1.weight"], beta_off=offsets[f"{p}.hh_gates.1.bias"])
            out[f"{tag}_LSTM{layer}"], out[f"{tag}_LSTM{layer}_HH"] = ih, hh
    for tag, prefix in (("PRIOR", "prior"), ("POST", "posterior")):
        wm, wl = offsets[f"{prefix}.mu_net.weight"], offsets[f"{prefix}.logvar_net.weight"]
        bm, bl = offsets[f"{prefix}.mu_net.bias"], offsets[f"{prefix}.logvar_net.bias"]
        rows, bias = [], []
        for zc in range(64):
            if zc < z:
                rows += [wm + zc * g * 9, wl + zc * g * 9]
                bias += [bm + zc, bl + zc]
            else:
                rows += [-1, -1]
                bias += [-1, -1]
        out[f"{tag}_GAUSS"] = dict(row_off=rows, col_off=cols([(g, g)], 9), bias_off=bias, flip=0, w_off=wm)
    # ConvTranspose2d(64, 4, 3, 1, 1): weight (in, out, kh, kw); as a conv the taps are flipped (pack.py)
    wt, bt = offsets["decoder.upc5.1.weight"], offsets["decoder.upc5.1.bias"]
    nout = sd_shapes["decoder.upc5.1.weight"][1]
    out["DEC_UPC5_1"] = dict(row_off=[wt + co * 9 for co in range(nout)] + [-1] * (16 - nout),
                             col_off=[ci * nout * 9 for ci in range(64)],
                             bias_off=[bt + co for co in range(nout)] + [-1] * (16 - nout), flip=1, w_off=wt)
    return out


RECON_KINDS = {"l1": 0, "dontcare_l1": 1, "mse": 2, "dontcare_mse": 3}  # cfg.reconstruction_loss (trainer.py:149-161)


class RacTrainLayer(C.Structure):
    _fields_ = [("row_off", C.c_void_p), ("col_off", C.c_void_p), ("bias_off", C.c_void_p),
                ("gamma_off", C.c_longlong), ("beta_off", C.c_longlong), ("rmean_off", C.c_longlong),
                ("rvar_off", C.c_longlong), ("w_off", C.c_longlong), ("flip", C.c_int),
                ("cnorm_gamma_off", C.c_longlong), ("cnorm_beta_off", C.c_longlong),
                ("grad_off", C.c_longlong), ("grad_count", C.c_longlong), ("w_count", C.c_longlong)]


class RacTrainConfig(C.Structure):
    _fields_ = [("batch", C.c_int), ("steps", C.c_int), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("adam_eps", C.c_float), ("kl_beta", C.c_float), ("robot_pixel_weight", C.c_float),
                ("recon_kind", C.c_int), ("zero_robot", C.c_int), ("n_params", C.c_longlong), ("n_buffers", C.c_longlong),
                ("fixed_skip", C.c_int)]


class RacTrainBatch(C.Structure):
    _fields_ = [("images", C.c_void_p), ("masks", C.c_void_p), ("states", C.c_void_p), ("actions", C.c_void_p),
                ("eps_prior", C.c_void_p), ("eps_post", C.c_void_p), ("seed", C.c_ulonglong), ("losses", C.c_void_p),
                ("true_token", C.c_void_p), ("noise_step", C.c_ulonglong), ("batch_weight", C.c_void_p),
                ("grads_ready", C.c_void_p), ("grads_ready_user", C.c_void_p), ("grads_ready_min", C.c_longlong),
                ("defer_unpack", C.c_int), ("defer_min", C.c_longlong)]


GRADS_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong)
OVERLAP_MIN_ELEMS = 4 << 20  # layers with at least this many weights are all-reduced while the backward pass runs
# single-tensor convolutions with at least this many weights take the fused optimizer step (RAC_FUSED_MIN_ELEMS: A/B switch)
FUSED_MIN_ELEMS = int(os.environ.get("RAC_FUSED_MIN_ELEMS", 1 << 20))


def _device_f32_view(ptr, count, device):
    """A float32 CUDA tensor over `count` elements at device address `ptr` (a buffer inside the library's arena)."""
    class _W:
        __cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_W(), device=device)


def complement_ranges(done, n):
    """Sorted disjoint (offset, count) ranges covering [0, n) minus the `done` ranges."""
    out, pos = [], 0
    for off, cnt in sorted(done):
        if off < pos:
            raise ValueError("overlapping gradient ranges")
        if off > pos:
            out.append((pos, off - pos))
        pos = off + cnt
    if pos > n:
        raise ValueError("gradient range beyond the buffer")
    if pos < n:
        out.append((pos, n - pos))
    return out


def adam_state_dict(m, v, steps_taken, layout, lr, beta1):
    """Flat first / second moments + one step count -> the dict torch.optim.Adam.state_dict() returns. `layout`:
    (offset, shape) per parameter in optimizer order. Like torch, no per-parameter state exists before the first step."""
    state = {}
    if steps_taken > 0:
        for i, (off, shape) in enumerate(layout):
            n = int(np.prod(shape)) if len(shape) else 1
            state[i] = {"step": torch.tensor(float(steps_taken)), "exp_avg": m[off:off + n].view(shape).clone(),
                        "exp_avg_sq": v[off:off + n].view(shape).clone()}
    group = {"lr": lr, "betas": (beta1, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "decoupled_weight_decay": False, "params": list(range(len(layout)))}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(state_dict, m, v, layout):
    """Inverse of adam_state_dict for a dict saved by the reference trainer (any torch version: "step" is an int in
    torch 1.x and a tensor later). Fills m / v in place; returns (steps_taken, lr, beta1)."""
    groups = state_dict["param_groups"]
    if len(groups) != 1:
        raise ValueError("loaded state dict has a different number of parameter groups")
    grp = groups[0]
    ids = list(grp["params"])
    if len(ids) != len(layout):
        raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
    if grp.get("amsgrad", False) or grp.get("weight_decay", 0) != 0 or grp.get("maximize", False):
        raise NotImplementedError("the B200 Adam kernel implements the reference's plain Adam (no amsgrad / weight decay)")
    if abs(float(grp.get("eps", 1e-8)) - 1e-8) > 0 or abs(float(grp["betas"][1]) - 0.999) > 0:
        raise NotImplementedError("eps / beta2 other than the reference's 1e-8 / 0.999")
    steps = set()
    m.zero_()
    v.zero_()
    for pos, pid in enumerate(ids):
        st = state_dict["state"].get(pid)
        if st is None:
            continue
        off, shape = layout[pos]
        if tuple(st["exp_avg"].shape) != tuple(shape):
            raise ValueError(f"optimizer state of parameter {pos} has shape {tuple(st['exp_avg'].shape)}, expected {tuple(shape)}")
        n = st["exp_avg"].numel()
        m[off:off + n].copy_(st["exp_avg"].reshape(-1))
        v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(st["step"]))
    if len(steps) > 1:
        raise ValueError(f"per-parameter Adam step counts differ ({sorted(steps)}): one flat update has one count")
    return (steps.pop() if steps else 0), float(grp["lr"]), float(grp["betas"][0])


class SVGTrainer:
    def __init__(self, config, model, process_group=None):
        self._config = config
        self.model = model
        c = svg_config_from(config)
        self.n_future = getattr(config, "n_future", 5)
        self.n_past = getattr(config, "n_past", 1)
        # n_past only sets the clip length (n_past + n_future frames, trainer.py:352). With last_frame_skip False (the
        # config default) the decoder keeps the skips of the clip's FIRST frame whatever n_past is: the model returns
        # the skip it used (dynamics.py:586-588,644), so `skip = curr_skip` (trainer.py:409-411) never changes it
        self._scheduled_sampling = bool(getattr(config, "scheduled_sampling", False))
        self._ss_k = float(getattr(config, "scheduled_sampling_k", 4000))
        self._forced_tokens = None
        kind = c.reconstruction_loss
        if kind not in RECON_KINDS:  # the reference raises the same for anything else (trainer.py:160-161)
            raise NotImplementedError(f"{kind}")
        self._fixed_skip = int(not c.last_frame_skip)
        self.process_group = process_group
        self.allreduce_events = None
        self.overlap_allreduce = True   # start the all-reduce of the large layers underneath the backward pass
        self.fused_update = True        # train_step: packed gradient -> Adam -> bf16 operand in one pass for the large layers
        self._deferred = False
        self._grad_scale_world = 1
        self._pending = []
        self._lib = _lib.load()
        dev = model._device
        # ---- flat fp32 storage; the module's parameters / running stats become views into it
        params = [(k, p) for k, p in model.named_parameters()]
        bufs = [(k, b) for k, b in model.named_buffers() if b.is_floating_point()]
        self._offsets, self._boffsets = {}, {}
        n = 0
        for k, p in params:
            self._offsets[k] = n
            n += p.numel()
        nb = 0
        for k, b in bufs:
            self._boffsets[k] = nb
            nb += b.numel()
        self.params = torch.empty(n, device=dev)
        self.buffers = torch.empty(max(nb, 1), device=dev)
        for k, p in params:
            o = self._offsets[k]
            view = self.params[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        for k, b in bufs:
            o = self._boffsets[k]
            view = self.buffers[o:o + b.numel()].view(b.shape)
            view.copy_(b.data)
            b.data = view
        self.grads = torch.zeros(n, device=dev)
        self.adam_m = torch.zeros(n, device=dev)
        self.adam_v = torch.zeros(n, device=dev)
        self.losses = torch.zeros(4, device=dev)
        # BatchNorm2d.num_batches_tracked counters: one flat int64 tensor (the module buffers are views), advanced by
        # one vector add per step -- the encoder runs twice per predicted frame (dynamics.py:566,619), the decoder once
        tracked = [(k, b) for k, b in model.named_buffers() if k.endswith("num_batches_tracked")]
        self._tracked = torch.zeros(max(len(tracked), 1), dtype=torch.long, device=dev)
        for i, (k, b) in enumerate(tracked):
            self._tracked[i] = b.to(dev)
            b.data = self._tracked[i]
        self._tracked_inc = torch.tensor([2 if k.startswith("encoder.") else 1 for k, _ in tracked] or [0],
                                         dtype=torch.long, device=dev)
        self._tables = _layer_tables(model, self._offsets, self._boffsets)
        self._keep = []  # device index arrays referenced by the library
        names = pack.LAYER_IDS + (pack.GN_LAYER_IDS if c.lstm_group_norm else [])
        layers = (RacTrainLayer * len(names))()
        for i, name in enumerate(names):
            t = self._tables[name]
            ro = torch.tensor(t["row_off"], dtype=torch.int64, device=dev)
            co = torch.tensor(t["col_off"], dtype=torch.int32, device=dev)
            self._keep += [ro, co]
            layers[i].row_off, layers[i].col_off = ro.data_ptr(), co.data_ptr()
            if t.get("bias_off") is not None:
                bo = torch.tensor(t["bias_off"], dtype=torch.int64, device=dev)
                self._keep.append(bo)
                layers[i].bias_off = bo.data_ptr()
            for f in ("gamma_off", "beta_off", "rmean_off", "rvar_off", "cnorm_gamma_off", "cnorm_beta_off"):
                setattr(layers[i], f, t.get(f, -1))
            layers[i].w_off = t["w_off"]
            layers[i].flip = t["flip"]
            layers[i].grad_off, layers[i].grad_count = t.get("grad_off", 0), t.get("grad_count", 0)
            layers[i].w_count = t.get("w_count", 0)
        self._layers = layers
        self._created_for = None
        self._lr = float(getattr(config, "lr", 1e-4))
        self._beta1 = float(getattr(config, "beta1", 0.9))
        self._kl_beta = float(getattr(config, "beta", 1e-4))
        self._rpw = float(getattr(config, "robot_pixel_weight", 0.0))
        self._kind = RECON_KINDS[kind]
        self._zero_robot = int("dontcare" in kind or bool(c.black_robot_input))
        # reparameterisation noise: Philox keyed on (seed, global training step, time step). Data-parallel replicas
        # seed alike (torch.initial_seed() is the same on every rank), so the rank is mixed in: every replica must
        # draw its own eps for its own batch
        rank = dist.get_rank(process_group) if process_group is not None else 0
        self._seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF
        self._eps = None
        self._step = 0
        self._adam_t = 0  # Adam steps taken (torch.optim.Adam state "step")
        model._ts = self  # the latest trainer owns the flat storage the model's train-mode forward works on

    # ---- scheduled sampling (reference trainer.py:132-147): same formula, same use of the global numpy generator
    def _schedule_prob(self):
        use_truth = self._ss_k / (self._ss_k + np.exp(self._step / self._ss_k))
        return [use_truth, 1 - use_truth]

    def _use_true_token(self):
        if not self._scheduled_sampling:
            return True
        return bool(np.random.choice([True, False], p=self._schedule_prob()))

    def set_true_tokens(self, tokens):
        """Test hook: per-step decisions (index 0 is ignored, the first frame is always ground truth) for the next step."""
        self._forced_tokens = tokens

    def set_noise(self, eps_prior, eps_post):
        """Test hook: (T-1, B, z_dim, 6, 8) reparameterisation noise for the next step (prior drawn first, lstm.py:276-279)."""
        self._eps = (eps_prior, eps_post)

    def _ensure(self, B, S):
        if self._created_for == (B, S):
            return
        cfg = RacTrainConfig(batch=B, steps=S, lr=self._lr, beta1=self._beta1, beta2=0.999, adam_eps=1e-8,
                             kl_beta=self._kl_beta, robot_pixel_weight=self._rpw, recon_kind=self._kind,
                             zero_robot=self._zero_robot, n_params=self.params.numel(), n_buffers=self.buffers.numel(),
                             fixed_skip=self._fixed_skip)
        m = self.model
        _lib.check(self._lib.rac_train_create(m.handle, C.byref(cfg), self._layers, _lib.ptr(self.params),
                                              _lib.ptr(self.buffers), _lib.ptr(self.grads), _lib.ptr(self.adam_m),
                                              _lib.ptr(self.adam_v)), m.handle, "rac_train_create")
        # the bias-correction step count survives a re-creation (new batch shape, new hyper-parameters, resume)
        _lib.check(self._lib.rac_train_set_adam_step(m.handle, self._adam_t), m.handle, "rac_train_set_adam_step")
        self._grad_scale_world = 1  # (a fresh training state scales by 1)
        self._created_for = (B, S)

    def forward_backward(self, batch, fused_update=False):
        """Fills self.grads and self.losses (sum over steps of recon, of KL). Returns the loss tensor.
        fused_update=True (what train_step does): optimizer_step() follows immediately and nobody reads the flat gradient
        of the large convolutions, so their gradient stays in the packed form the wgrad GEMM wrote and the optimizer
        updates those layers in one pass (rac_train_batch.defer_unpack); `unpack_deferred()` materialises it after all."""
        m = self.model
        dev = m._device
        f32 = lambda t: None if t is None else t.to(device=dev, dtype=torch.float32).contiguous()
        images, actions = f32(batch["images"]), f32(batch["actions"])
        masks, states = f32(batch.get("masks")), f32(batch.get("states"))
        # the reference unrolls range(1, n_past + n_future) whatever the clip length is (trainer.py:352) and uses the
        # first min(batch_size, B) clips only implicitly (init_hidden, :349-350); a shorter clip is an IndexError there
        T_ref = self.n_past + self.n_future
        if images.shape[0] < T_ref:
            raise IndexError(f"clip of {images.shape[0]} frames, n_past + n_future = {T_ref} (reference trainer.py:352)")
        if images.shape[0] > T_ref:
            images, actions = images[:T_ref].contiguous(), actions[:T_ref - 1].contiguous()
            masks = None if masks is None else masks[:T_ref].contiguous()
            states = None if states is None else states[:T_ref].contiguous()
        T, B = images.shape[0], images.shape[1]
        if B % 4:
            raise ValueError(f"batch of {B} clips: the B200 training step needs a multiple of 4 (64-row GEMM tiles on the "
                             "6x8 latent map); drop or pad the trailing partial batch (DataLoader(drop_last=True))")
        bw = None
        if getattr(self._config, "load_movement_info", False):  # trainer.py:426-429
            info = batch["high_movement"].to(dev).bool()
            bw = (float(self._config.movement_weight) * info).float()
            bw[~info] = 1.0
            bw = bw.contiguous()
        self._ensure(B, T - 1)
        eps_p = eps_q = None
        if self._eps is not None:
            eps_p, eps_q = f32(self._eps[0]), f32(self._eps[1])
            self._eps = None
        if self._forced_tokens is not None:
            decisions, self._forced_tokens = [bool(x) for x in self._forced_tokens], None
        else:  # one draw per step i > 1, in step order, exactly as the reference loop (trainer.py:352-356)
            decisions = [True] + [self._use_true_token() for _ in range(1, T - 1)]
        tokens = np.ascontiguousarray(np.array([1] + [int(d) for d in decisions[1:T - 1]], dtype=np.int32))
        bt = RacTrainBatch(images=_lib.ptr(images), masks=_lib.ptr(masks), states=_lib.ptr(states),
                           actions=_lib.ptr(actions), eps_prior=_lib.ptr(eps_p), eps_post=_lib.ptr(eps_q),
                           seed=self._seed, losses=_lib.ptr(self.losses),
                           true_token=None if all(tokens) else tokens.ctypes.data, noise_step=self._step,
                           batch_weight=_lib.ptr(bw))
        self._pending = []  # (flat offset, flat count, Work, tensor) of the all-reduces started underneath the backward pass
        dp = self._dp_world() > 1
        if dp and self.overlap_allreduce:
            def ready(_user, ptr, count, flat_off, flat_count):
                # runs on this thread while the library is still enqueueing: NCCL orders the collective after
                # everything enqueued on the current stream so far and runs it on its own stream. `ptr` is either a
                # range of the flat gradient buffer or (fused optimizer step) a layer's packed gradient buffer.
                t = _device_f32_view(ptr, count, dev)
                w = dist.all_reduce(t, group=self.process_group, async_op=True)
                self._pending.append((int(flat_off), int(flat_count), w, t))
            self._ready_cb = GRADS_READY_FN(ready)  # (kept alive for the duration of the call)
            bt.grads_ready = C.cast(self._ready_cb, C.c_void_p)
            bt.grads_ready_min = OVERLAP_MIN_ELEMS
        # (data parallel without the overlap has no hook through which the packed gradients could be reduced)
        self._deferred = bool(fused_update and self.fused_update and (not dp or self.overlap_allreduce))
        bt.defer_unpack = int(self._deferred)
        bt.defer_min = FUSED_MIN_ELEMS
        _lib.check(self._lib.rac_train_forward_backward(m.handle, C.byref(bt), _lib.stream_ptr()), m.handle,
                   "rac_train_forward_backward")
        self._keep_batch = (images, actions, masks, states, eps_p, eps_q, bw)  # alive until the stream has consumed them
        return self.losses

    def _dp_world(self):
        return dist.get_world_size(self.process_group) if self.process_group is not None else 1

    def optimizer_step(self):
        m = self.model
        world = self._dp_world()
        if world > 1:
            ev = self.allreduce_events  # optional (start, end) CUDA events: bench.py reports the EXPOSED all-reduce time
            if ev is not None:
                ev[0].record()
            # the large layers were handed to NCCL while BPTT was still running (forward_backward); what is left is the
            # complement of their ranges. The mean is taken inside the Adam kernel (gradient scale 1 / world).
            pending = getattr(self, "_pending", [])
            for off, cnt in complement_ranges([(p[0], p[1]) for p in pending], self.grads.numel()):
                dist.all_reduce(self.grads[off:off + cnt], group=self.process_group)
            for p in pending:
                p[2].wait()
            self._pending = []
            if ev is not None:
                ev[1].record()
        if world != self._grad_scale_world:
            _lib.check(self._lib.rac_train_set_grad_scale(m.handle, 1.0 / world), m.handle, "rac_train_set_grad_scale")
            self._grad_scale_world = world
        _lib.check(self._lib.rac_train_adam_step(m.handle, _lib.stream_ptr()), m.handle, "rac_train_adam_step")
        m._packed_dirty = True  # the eval-mode packed copy (folded BatchNorm) is stale now
        self._tracked.add_(self._tracked_inc, alpha=self._created_for[1])  # num_batches_tracked += forwards this step
        self._adam_t += 1
        self._step += 1

    def train_step(self, batch):
        """One reference `_train_step`: returns {"recon_loss", "kld"} averaged over n_future (trainer.py:463-464)."""
        if not self.model.training:
            raise RuntimeError("call model.train() before train_step (reference trainer.py:754)")
        losses = self.forward_backward(batch, fused_update=True)
        self.optimizer_step()
        # the reference divides the logged sums by cfg.n_future whatever the clip length is (trainer.py:463-464)
        nf = int(getattr(self._config, "n_future", batch["images"].shape[0] - 1))
        vals = losses.cpu()
        out = {"recon_loss": float(vals[0]) / nf, "kld": float(vals[1]) / nf}
        if batch.get("masks") is not None:  # the reference logs these for every step (trainer.py:436-439)
            out["robot_loss"], out["world_loss"] = float(vals[2]) / nf, float(vals[3]) / nf
        return out

    @torch.no_grad()
    def _eval_step(self, data, autoregressive=False):
        """PredictionTrainer._eval_step (trainer.py:566-734) for the svg model, single view: evaluates a snippet of
        cfg.n_eval frames with the eval-mode model (prior z, posterior only for the KL), returning the reference's
        dict of averaged metrics ("1step_*" / "autoreg_*", plus "<i>_step_*" for autoregressive rollouts).
        data: images (T,B,3,H,W), states, actions, masks (true masks, metrics), pred_masks (model inputs)."""
        from . import losses as L
        from . import metrics as M
        from .image import zero_robot_region

        m = self.model
        if m.training:
            raise RuntimeError("call model.eval() before _eval_step (reference trainer.py:785-790)")
        c = m._c
        dev = m._device
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        x, states, ac = f32(data["images"]), f32(data["states"]), f32(data["actions"])
        true_masks, masks = f32(data["masks"]), f32(data["pred_masks"])
        n_eval = int(getattr(self._config, "n_eval", x.shape[0]))
        bs = min(int(getattr(self._config, "test_batch_size", x.shape[1])), x.shape[1])
        if bs != x.shape[1]:
            raise ValueError("test_batch_size smaller than the batch: the reference would fail in forward as well")
        m.init_hidden(bs)
        eps = self._eps
        self._eps = None
        prefix = "autoreg" if autoregressive else "1step"
        acc = torch.zeros(6, device=dev)  # recon, robot, world, psnr, ssim, kld: summed on device, one read at the end
        per_step = []
        dontcare = ("dontcare" in c.reconstruction_loss) or bool(c.black_robot_input)
        x_pred = skip = None
        for i in range(1, n_eval):
            x_j = x_pred if (autoregressive and i > 1) else x[i - 1]
            m_j, r_j, a_j = masks[i - 1], states[i - 1], ac[i - 1]
            x_i, m_i, r_i = x[i], masks[i], states[i]
            x_j_black = zero_robot_region(m_j, x_j) if dontcare else x_j
            if c.last_frame_skip:
                skip = None
            m_in = torch.cat([m_j, m_i], 1) if c.model_use_future_mask else m_j
            r_in = (r_j, r_i) if c.model_use_future_robot_state else r_j
            if eps is not None:
                m.set_noise(eps=eps[0][i - 1], eps_post=eps[1][i - 1])
            x4, curr_skip, mu, logvar, mu_p, logvar_p = m.forward(
                x_j_black, m_in if c.model_use_mask else None, r_in if c.model_use_robot_state else None, None, a_j,
                x_i, None, r_i if c.model_use_robot_state else None, None, skip, force_use_prior=True)
            x_pred = torch.empty_like(x_j)
            _lib.check(self._lib.rac_composite(_lib.ptr(x4), _lib.ptr(x_j), _lib.ptr(x_pred), bs, 48 * 64,
                                               _lib.stream_ptr()), m.handle, "rac_composite")
            if i <= self.n_past:
                skip = curr_skip
            tm = true_masks[i]
            if c.reconstruction_loss == "l1":
                recon = L.l1_criterion(x_pred, x_i)
            else:
                recon = L.dontcare_l1_criterion(x_pred, x_i, tm, self._rpw)
            rw = L._robot_world_mse(x_pred, x_i, tm)
            p = M.psnr(x_i, x_pred, mask=tm, clamp01=True).mean()
            s = M.ssim_mean(x_i, x_pred, mask=tm)
            kl = L.kl_criterion(mu, logvar, mu_p, logvar_p, bs)
            step = torch.stack([recon, rw[0], rw[1], p, s, kl])
            acc += step
            per_step.append(step)
        vals = (acc / (n_eval - 1)).cpu().tolist()
        out = {f"{prefix}_{k}": v for k, v in zip(("recon_loss", "robot_loss", "world_loss", "psnr", "ssim", "kld"), vals)}
        if autoregressive:
            for i, st in enumerate(torch.stack(per_step).cpu().tolist(), start=1):
                out[f"{i}_step_psnr"], out[f"{i}_step_ssim"], out[f"{i}_step_world_loss"] = st[3], st[4], st[2]
        return out

    eval_step = _eval_step

    def _eval_video(self, data, autoregressive=False, noise=None):
        """PredictionTrainer._eval_video (trainer.py:489-565): the video is cut into floor(T / n_eval) windows, every
        window is evaluated with `_eval_step` (3 samples for autoregressive svg evaluation of a "finetune" experiment,
        else 1), the samples are ranked by "autoreg_psnr" and the best one is averaged over the windows.
        `data["pred_masks"]` (model-input masks, e.g. from the analytical robot model) is used when present, else the
        true masks -- the reference obtains them from `self.robot_model.predict_batch` inside this function, which
        needs the simulator and stays outside this package. `noise[sample][window]` = (eps_prior, eps_post) is a test
        hook for `set_noise`."""
        cf = self._config
        num_samples = 3 if (autoregressive and "finetune" in str(getattr(cf, "experiment", ""))) else 1
        x = data["images"]
        T = len(x)
        window = int(getattr(cf, "n_eval", T))
        nwin = T // window
        sampled = [dict() for _ in range(num_samples)]
        pm = data.get("pred_masks", data["masks"])
        for i in range(nwin):
            s, e = i * window, (i + 1) * window
            batch = {"images": x[s:e], "states": data["states"][s:e], "actions": data["actions"][s:e - 1],
                     "masks": data["masks"][s:e], "pred_masks": pm[s:e]}
            for k in range(num_samples):
                if noise is not None:
                    self.set_noise(*noise[k][i])
                for key, v in self._eval_step(batch, autoregressive).items():
                    sampled[k][key] = sampled[k].get(key, 0.0) + v
        if autoregressive:
            sampled.sort(key=lambda d: d["autoreg_psnr"], reverse=True)
        return {k: v / nwin for k, v in sampled[0].items()}

    # ---- checkpoints (reference trainer.py:829-896): {"model", "optimizer", "step"} with torch.optim.Adam's state_dict
    def _adam_layout(self):
        return [(self._offsets[k], tuple(p.shape)) for k, p in self.model.named_parameters()]

    def optimizer_state_dict(self):
        """torch.optim.Adam(self.model.parameters(), lr, (beta1, 0.999)).state_dict() of the reference trainer
        (trainer.py:109-122): parameter i of the group is the i-th entry of model.parameters(), same order as the
        reference model's."""
        return adam_state_dict(self.adam_m, self.adam_v, self._adam_t, self._adam_layout(), self._lr, self._beta1)

    def load_optimizer_state_dict(self, state_dict):
        t, lr, beta1 = load_adam_state_dict(state_dict, self.adam_m, self.adam_v, self._adam_layout())
        self._adam_t, self._lr, self._beta1 = t, lr, beta1
        self._created_for = None  # hyper-parameters and the step count are handed over at the next rac_train_create

    def save_checkpoint(self, path):
        """PredictionTrainer._save_checkpoint (trainer.py:829-837): the same three keys. Data parallel: BatchNorm running
        statistics are per rank (as in per-process training of the reference); rank 0's are broadcast first so that every
        rank holds -- and any rank may write -- the same checkpoint."""
        if self.process_group is not None and dist.get_world_size(self.process_group) > 1:
            dist.broadcast(self.buffers, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
        torch.save({"model": self.model.state_dict(), "optimizer": self.optimizer_state_dict(), "step": self._step}, path)

    def load_checkpoint(self, path):
        """PredictionTrainer._load_checkpoint with a given path (trainer.py:884-896): loads the model; a "finetune"
        experiment restarts at step 0 with a fresh optimizer, anything else resumes step and optimizer. Returns the step."""
        ckpt = torch.load(path, map_location=self.model._device)
        self.model.load_state_dict(ckpt["model"])
        if "finetune" in str(getattr(self._config, "experiment", "")):
            self._step = 0
        else:
            self._step = int(ckpt["step"])
            self.load_optimizer_state_dict(ckpt["optimizer"])
        return self._step

    def unpack_deferred(self):
        """After forward_backward(fused_update=True): write the flat gradient of the layers whose gradient stayed packed
        (inspection; optimizer_step does not need it)."""
        m = self.model
        _lib.check(self._lib.rac_train_unpack_deferred(m.handle, _lib.stream_ptr()), m.handle, "rac_train_unpack_deferred")

    def invalidate_packed(self):
        """Parameters were changed from outside (load_state_dict, an external optimizer): re-pack at the next step."""
        m = self.model
        _lib.check(self._lib.rac_train_invalidate_packed(m.handle), m.handle, "rac_train_invalidate_packed")

    def _publish_grads(self):
        """Train-mode SVGConvModel.forward under autograd: after the backward pass of step 0 every parameter's .grad
        is its slice of the flat gradient buffer (what loss.backward() leaves behind for torch.optim.Adam)."""
        for k, p in self.model.named_parameters():
            o = self._offsets[k]
            g = self.grads[o:o + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() == g.data_ptr():
                p.grad = g
            else:  # gradients accumulate across backward() calls until zero_grad(), as in torch
                p.grad = p.grad + g

    def grad_of(self, key):
        o = self._offsets[key]
        p = dict(self.model.named_parameters())[key]
        return self.grads[o:o + p.numel()].view(p.shape)
