"""Drop-in for the reference `SVGConvModel` (src/prediction/models/dynamics.py:457-644) on B200.

Same constructor (`SVGConvModel(cfg)`), same `init_hidden(batch_size)`, same `forward(...)` signature and return
tuple, same `state_dict()` keys -- a checkpoint saved by the reference trainer (`{"model": state_dict, ...}`,
trainer.py:829-837) loads with `load_state_dict`. The arithmetic is done by libracb200.so: eval-mode BatchNorm folded
into bf16 weights, tcgen05 implicit-GEMM convolutions with fused LSTM / z-sample / sigmoid epilogues.

Train mode (`model.train()`, the nn.Module default): `init_hidden` + `forward` run the training tape of
libracb200.so one step per call inside the caller's torch autograd graph (`_TrainStep`), so the reference's
`_train_step` body (trainer.py:326-465 -- compositing, `_recon_loss`, `kl_criterion`, `loss.backward()`,
`torch.optim.Adam(model.parameters())`) runs unchanged on this class. `SVGTrainer.train_step` is the fused, faster
way to do the same step (one C call for forward + BPTT, flat Adam).

Not supported (raises): heatmaps; train-mode forward without `next_image` (the reference never trains on the prior's z).
lstm_group_norm=True (NormConvLSTMCell, lstm.py:151-198) is supported.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, pack
from .config import svg_config_from, validate_model_config


def _spec(c):
    """state_dict key -> (shape, kind) in the reference's registration order (dynamics.py:467-516)."""
    g, z, a, r = c.g_dim, c.z_dim, c.action_dim, c.robot_dim
    spec = OrderedDict()

    def vgg(prefix, cin, cout):
        spec[f"{prefix}.main.0.weight"] = ((cout, cin, 3, 3), "conv_w")
        spec[f"{prefix}.main.1.weight"] = ((cout,), "bn_w")
        spec[f"{prefix}.main.1.bias"] = ((cout,), "zero")
        spec[f"{prefix}.main.1.running_mean"] = ((cout,), "buf_zero")
        spec[f"{prefix}.main.1.running_var"] = ((cout,), "buf_one")
        spec[f"{prefix}.main.1.num_batches_tracked"] = ((), "buf_long")

    def conv(prefix, cin, cout, k=3):
        spec[f"{prefix}.weight"] = ((cout, cin, k, k), "conv_w")
        spec[f"{prefix}.bias"] = ((cout,), "zero")

    def lstm_cell(prefix, k):
        if c.lstm_group_norm:  # NormConvLSTMCell (lstm.py:151-175); GroupNorm is not touched by init_weights
            for gk in ("ih_gates", "hh_gates"):
                conv(f"{prefix}.{gk}.0", g, 4 * g, k)
                spec[f"{prefix}.{gk}.1.weight"] = ((4 * g,), "one")
                spec[f"{prefix}.{gk}.1.bias"] = ((4 * g,), "zero")
            spec[f"{prefix}.c_norm.weight"] = ((g,), "one")
            spec[f"{prefix}.c_norm.bias"] = ((g,), "zero")
        else:
            conv(f"{prefix}.gates", 2 * g, 4 * g, k)

    nc = c.channels + (1 if c.model_use_mask else 0) + (1 if (c.model_use_mask and c.model_use_future_mask) else 0)
    for name, cin, cout in [("c1.0", nc, 64), ("c1.1", 64, 64), ("c2.0", 64, 128), ("c2.1", 128, 128),
                            ("c3.0", 128, 256), ("c3.1", 256, 256), ("c3.2", 256, 256), ("c4.0", 256, 512),
                            ("c4.1", 512, 512), ("c4.2", 512, g)]:
        vgg(f"encoder.{name}", cin, cout)
    extra = (r if c.model_use_robot_state else 0)
    extra2 = (r if c.model_use_future_robot_state else 0)
    conv("frame_pred_input_conv", g + a + z + extra + extra2, g)
    for layer, k in ((0, 5), (1, 3)):
        lstm_cell(f"frame_predictor.lstm.{layer}", k)
    conv("posterior_input_conv", g + extra, g)
    conv("prior_input_conv", g + a + extra + extra2, g)
    for p in ("posterior", "prior"):
        for layer, k in ((0, 5), (1, 3)):
            lstm_cell(f"{p}.lstm.{layer}", k)
        conv(f"{p}.mu_net", g, z)
        conv(f"{p}.logvar_net", g, z)
    for name, cin, cout in [("upc2.0", g, 512), ("upc2.1", 512, 512), ("upc2.2", 512, 256), ("upc3.0", 512, 256),
                            ("upc3.1", 256, 256), ("upc3.2", 256, 128), ("upc4.0", 256, 128), ("upc4.1", 128, 64),
                            ("upc5.0", 128, 64)]:
        vgg(f"decoder.{name}", cin, cout)
    spec["decoder.upc5.1.weight"] = ((64, c.channels + 1, 3, 3), "conv_w")
    spec["decoder.upc5.1.bias"] = ((c.channels + 1,), "zero")
    return spec


class _TrainSkips:
    """`skip` returned by a train-mode forward: the skips live on the library's tape (their gradient flows there, not
    through torch); handing the object back as `skip=` selects the first frame's skips (last_frame_skip False)."""

    def __init__(self, epoch):
        self.epoch = epoch


class _TrainStep(torch.autograd.Function):
    """One time step of the SVG model on the library's training tape. The hidden state and every intermediate stay
    inside libracb200.so; `token` chains the steps so that autograd runs their backward passes in reverse time order
    (BPTT through the ConvLSTM states happens inside the library)."""

    @staticmethod
    def forward(ctx, model, token, image, mask, robot, next_robot, action, eps_p, eps_q, keep_skip):
        ts = model._ts
        n, z = image.shape[0], model._c.z_dim
        dev = image.device
        x_pred = torch.empty(n, 4, 48, 64, device=dev)
        mu, logvar, mu_p, logvar_p = (torch.empty(n, z, 6, 8, device=dev) for _ in range(4))
        step = _lib.RacTrainStep(
            image=_lib.ptr(image), mask=_lib.ptr(mask), robot=_lib.ptr(robot), next_robot=_lib.ptr(next_robot),
            action=_lib.ptr(action), eps_prior=_lib.ptr(eps_p), eps_post=_lib.ptr(eps_q), seed=ts._seed,
            noise_step=model._noise_ctr, keep_skip=int(keep_skip), x_pred=_lib.ptr(x_pred), mu=_lib.ptr(mu),
            logvar=_lib.ptr(logvar), mu_p=_lib.ptr(mu_p), logvar_p=_lib.ptr(logvar_p))
        _lib.check(model._lib.rac_train_step_forward(model._h, C.byref(step), _lib.stream_ptr()), model._h,
                   "rac_train_step_forward")
        ctx.model, ctx.t, ctx.epoch, ctx.keep_skip = model, model._train_t, model._train_epoch, int(keep_skip)
        ctx.inputs = (image, mask, robot, next_robot, action)  # inputs only (no graph cycle): alive until backward
        return x_pred, mu, logvar, mu_p, logvar_p, token.new_zeros(())

    @staticmethod
    def backward(ctx, dx, dmu, dlv, dmu_p, dlv_p, _dtoken):
        model = ctx.model
        if ctx.epoch != model._train_epoch:
            raise RuntimeError("backward through a train-mode forward of an earlier init_hidden(): the library keeps "
                               "one tape (retain_graph / second-order use is not supported)")
        image, mask, robot, next_robot, action = ctx.inputs
        f = lambda g: None if g is None else g.to(torch.float32).contiguous()
        dx, dmu, dlv, dmu_p, dlv_p = f(dx), f(dmu), f(dlv), f(dmu_p), f(dlv_p)
        dimage = torch.empty_like(image) if ctx.needs_input_grad[2] else None
        step = _lib.RacTrainStep(image=_lib.ptr(image), mask=_lib.ptr(mask), robot=_lib.ptr(robot),
                                 next_robot=_lib.ptr(next_robot), action=_lib.ptr(action), keep_skip=ctx.keep_skip)
        _lib.check(model._lib.rac_train_step_backward(model._h, ctx.t, C.byref(step), _lib.ptr(dx), _lib.ptr(dmu),
                                                      _lib.ptr(dlv), _lib.ptr(dmu_p), _lib.ptr(dlv_p),
                                                      _lib.ptr(dimage), _lib.stream_ptr()), model._h,
                   "rac_train_step_backward")
        ctx.inputs = None
        if ctx.t == 0:
            model._ts._publish_grads()
        return None, torch.zeros_like(_dtoken), dimage, None, None, None, None, None, None, None


class _Holder(nn.Module):
    """Parameter container; the module tree only exists to give state_dict() the reference's dotted keys."""


class _LazySkips:
    """The `skip` element of forward()'s return value: [h1, h2, h3, h4] as NCHW fp32 tensors, materialised from the
    NHWC bf16 workspace only when indexed (the planner never touches them)."""

    def __init__(self, model, batch):
        self._m, self._b, self._cache = model, batch, None

    def _load(self):
        if self._cache is None:
            m, b = self._m, self._b
            cat5 = m._buffer_view("cat5", (b, 48, 64, 128))[..., 64:]
            cat4 = m._buffer_view("cat4", (b, 24, 32, 256))[..., 128:]
            cat3 = m._buffer_view("cat3", (b, 12, 16, 512))[..., 256:]
            h4 = m._buffer_view("h4", (b, 6, 8, m._c.g_dim))
            self._cache = [t.permute(0, 3, 1, 2).float().contiguous() for t in (cat5, cat4, cat3, h4)]
        return self._cache

    def __getitem__(self, i):
        return self._load()[i]

    def __len__(self):
        return 4

    def __iter__(self):
        return iter(self._load())


class SVGConvModel(nn.Module):
    """Conv SVG LSTM predictor (reference dynamics.py:457)."""

    def __init__(self, config, conv_impl=None):
        super().__init__()
        self._config = config
        self._c = c = svg_config_from(config)
        validate_model_config(c)
        self._image_width, self._image_height = c.image_width, c.image_height
        if not torch.cuda.is_available():
            raise RuntimeError("SVGConvModel (B200) needs a CUDA device: there is no CPU path")
        self._device = torch.device("cuda", torch.cuda.current_device())
        # parameters under the reference's names; init = reference init_weights (base.py:26-36)
        for key, (shape, kind) in _spec(c).items():
            if kind == "conv_w":
                t = torch.empty(shape).normal_(0.0, 0.02)
            elif kind == "bn_w":
                t = torch.empty(shape).normal_(1.0, 0.02)
            elif kind in ("buf_one", "one"):
                t = torch.ones(shape)
            elif kind == "buf_long":
                t = torch.tensor(0, dtype=torch.long)
            else:
                t = torch.zeros(shape)
            self._register(key, t, buffer=kind.startswith("buf"))
        import os

        if conv_impl is None:
            conv_impl = 1 if os.environ.get("RAC_CONV_IMPL", "tc") == "simt" else 0
        self._lib = _lib.load()
        rc = _lib.RacConfig(c.image_height, c.image_width, c.g_dim, c.z_dim, c.action_dim, c.robot_dim,
                            int(bool(c.model_use_mask)), int(bool(c.model_use_mask and c.model_use_future_mask)),
                            int(bool(c.model_use_robot_state)),
                            int(bool(c.model_use_robot_state and c.model_use_future_robot_state)), int(conv_impl),
                            int(bool(c.lstm_group_norm)))
        h = C.c_void_p()
        code = self._lib.rac_create(C.byref(rc), C.byref(h))
        self._h = h
        _lib.check(code, h, "rac_create")
        self._packed_dirty = True
        self._batch = 0
        self._noise_ctr = 0
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._eps = None
        self._eps_post = None
        self._ts = None           # SVGTrainer owning the flat parameter / gradient storage (train-mode forward)
        self._train_epoch = 0     # init_hidden() calls in train mode: one library tape each
        self._train_t = 0
        self._train_token = None
        self.train(True)  # nn.Module default, as the reference; planning callers call .eval()

    # ------------------------------------------------------------------ parameter plumbing
    def _register(self, key, tensor, buffer):
        parts = key.split(".")
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Holder())
            mod = mod._modules[p]
        if buffer:
            mod.register_buffer(parts[-1], tensor)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._params_touched()
        return out

    def _apply(self, fn, *a, **k):
        # .to()/.cuda()/.float(): the packed copy inside the library is rebuilt from whatever the tensors become
        out = super()._apply(fn, *a, **k)
        self._params_touched()
        return out

    def _params_touched(self):
        """The parameter tensors changed behind the library's back: both packed copies are stale -- the eval-mode one
        (folded BatchNorm) and the training operands the fused optimizer step keeps across steps."""
        self._packed_dirty = True
        if getattr(self, "_ts", None) is not None and getattr(self, "_h", None):
            self._ts.invalidate_packed()

    def _ensure_packed(self):
        if not self._packed_dirty:
            return
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        for name, (w, b) in pack.pack_state_dict(sd, self._c).items():
            w = w.contiguous()
            b = b.contiguous()
            _lib.check(self._lib.rac_load_layer(self._h, pack.LAYER_INDEX[name], _lib.ptr(w), w.numel(), _lib.ptr(b),
                                                b.numel()), self._h, f"rac_load_layer({name})")
        if self._c.lstm_group_norm:
            for name, t in pack.pack_lstm_norm(sd, self._c).items():
                _lib.check(self._lib.rac_load_lstm_norm(self._h, pack.LAYER_INDEX[name], _lib.ptr(t), t.numel()),
                           self._h, f"rac_load_lstm_norm({name})")
        self._packed_dirty = False

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.rac_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _buffer_view(self, name, shape, dtype=torch.bfloat16):
        p, n, eb = C.c_void_p(), C.c_int64(), C.c_int()
        _lib.check(self._lib.rac_debug_buffer(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(eb)), self._h,
                   "rac_debug_buffer")
        numel = 1
        for s in shape:
            numel *= s
        assert numel == n.value, (name, shape, n.value)
        esize = torch.empty((), dtype=dtype).element_size()
        assert esize == eb.value
        iface = {"shape": tuple(shape), "typestr": {2: "<i2", 4: "<f4"}[esize], "data": (p.value, False), "version": 3}

        class _W:
            __cuda_array_interface__ = iface

        t = torch.as_tensor(_W(), device=self._device)
        return t.view(dtype) if dtype == torch.bfloat16 else t

    def prepare(self, batch_size):
        """Pack weights (if changed) and size the workspace for `batch_size` candidates."""
        self._ensure_packed()
        if batch_size != self._batch:
            _lib.check(self._lib.rac_prepare(self._h, int(batch_size)), self._h, "rac_prepare")
            self._batch = int(batch_size)

    def set_noise(self, eps=None, eps_post=None):
        """Test hook: the next forward() uses these (n, z_dim, 6, 8) tensors instead of Philox noise for the prior /
        posterior reparameterisation (reference draws them inside reparameterize, lstm.py:276-279)."""
        self._eps, self._eps_post = eps, eps_post

    def launch_count(self):
        return int(self._lib.rac_launch_count(self._h))

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ reference interface
    def init_hidden(self, batch_size=None):
        """dynamics.py:536-542."""
        if batch_size is None:
            batch_size = getattr(self._config, "batch_size", 16)
        if self.training:
            return self._train_begin(int(batch_size))
        self.prepare(batch_size)
        _lib.check(self._lib.rac_init_hidden(self._h, int(batch_size), _lib.stream_ptr()), self._h, "rac_init_hidden")

    def _f32(self, t):
        if t is None:
            return None
        return t.to(device=self._device, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ train mode (autograd step API)
    def _train_begin(self, batch_size):
        """init_hidden() in train mode: zero recurrent state = an empty tape; the bf16 operands are re-packed from the
        current parameters (torch.optim.Adam has just stepped them) and the flat gradient buffer is zeroed."""
        if self._ts is None:
            from .trainer import SVGTrainer

            SVGTrainer(self._config, self)  # flattens the parameters into one buffer and registers itself as _ts
        ts = self._ts
        steps = int(getattr(self._config, "n_past", 1)) + int(getattr(self._config, "n_future", 5)) - 1
        ts._ensure(batch_size, steps)
        _lib.check(self._lib.rac_train_step_begin(self._h, _lib.stream_ptr()), self._h, "rac_train_step_begin")
        self._train_epoch += 1
        self._train_t = 0
        self._train_batch = batch_size
        self._train_token = torch.zeros((), device=self._device, requires_grad=True)

    def _train_forward(self, image, mask, robot, action, next_image, next_robot, skip, force_use_prior, sample_mean):
        c = self._c
        if next_image is None or force_use_prior or sample_mean:
            raise NotImplementedError("train-mode forward is the training step of the reference (posterior z, "
                                      "trainer.py:384-399): next_image is required; call .eval() for prior rollouts")
        if self._train_token is None or image.shape[0] != self._train_batch:
            raise RuntimeError(f"init_hidden({image.shape[0]}) must be called (in train mode) before a train-mode forward")
        image, action = self._f32(image), self._f32(action)
        mask = self._f32(mask) if c.model_use_mask else None
        r = nr = None
        if c.model_use_robot_state:
            if c.model_use_future_robot_state:
                r, r2 = self._f32(robot[0]), self._f32(robot[1])
                nr = self._f32(next_robot)
                if r2.data_ptr() != nr.data_ptr() and not torch.equal(r2, nr):
                    raise NotImplementedError("train-mode forward: robot[1] and next_robot must be the same state (the "
                                              "reference trainer passes r_i for both, trainer.py:376-399)")
            else:
                r, nr = self._f32(robot), self._f32(next_robot)
        keep_skip = 0
        if skip is not None and not c.last_frame_skip:
            if not isinstance(skip, _TrainSkips) or skip.epoch != self._train_epoch or self._train_t == 0:
                raise NotImplementedError("train-mode forward: `skip` must be the object returned by an earlier forward "
                                          "of the same clip (the first frame's skips, trainer.py:370-371,409-411)")
            keep_skip = 1
        eps, eps_post = self._f32(self._eps), self._f32(self._eps_post)
        out = _TrainStep.apply(self, self._train_token, image, mask, r, nr, action, eps, eps_post, keep_skip)
        x_pred, mu, logvar, mu_p, logvar_p, self._train_token = out
        self._train_t += 1
        self._noise_ctr += 1
        self._eps = self._eps_post = None
        self._packed_dirty = True  # the eval-mode packed copy (folded BatchNorm) is stale: statistics moved
        self._ts._tracked.add_(self._ts._tracked_inc)  # BatchNorm2d.num_batches_tracked (encoder runs twice, :566,619)
        return x_pred, _TrainSkips(self._train_epoch), mu, logvar, mu_p, logvar_p

    def forward(self, image, mask, robot, heatmap, action, next_image=None, next_mask=None, next_robot=None,
                next_heatmap=None, skip=None, force_use_prior=False, sample_mean=False):
        """dynamics.py:544-644. Returns (x_pred, skip, mu, logvar, mu_p, logvar_p)."""
        if self.training:
            return self._train_forward(image, mask, robot, action, next_image, next_robot, skip, force_use_prior,
                                       sample_mean)
        with torch.no_grad():
            return self._eval_forward(image, mask, robot, heatmap, action, next_image, next_mask, next_robot,
                                      next_heatmap, skip, force_use_prior, sample_mean)

    def _eval_forward(self, image, mask, robot, heatmap, action, next_image=None, next_mask=None, next_robot=None,
                      next_heatmap=None, skip=None, force_use_prior=False, sample_mean=False):
        c = self._c
        n = image.shape[0]
        if self._batch != n:
            raise RuntimeError(f"init_hidden({n}) must be called before forward on a batch of {n} "
                               f"(hidden state is sized for {self._batch})")
        image = self._f32(image)
        action = self._f32(action)
        mask = self._f32(mask) if c.model_use_mask else None
        r = r2 = None
        if c.model_use_robot_state:
            if c.model_use_future_robot_state:
                r, r2 = self._f32(robot[0]), self._f32(robot[1])
            else:
                r = self._f32(robot)
        use_post = next_image is not None
        keep_skip = 0
        if skip is not None and not c.last_frame_skip:
            keep_skip = 1
            if not isinstance(skip, _LazySkips):
                self._write_skips(skip, n)
        z = c.z_dim
        x_pred = torch.empty(n, 4, 48, 64, device=self._device)
        mu_p = torch.empty(n, z, 6, 8, device=self._device)
        logvar_p = torch.empty_like(mu_p)
        mu = torch.empty_like(mu_p) if use_post else None
        logvar = torch.empty_like(mu_p) if use_post else None
        eps = self._f32(self._eps)
        eps_post = self._f32(self._eps_post)
        nr = self._f32(next_robot) if (use_post and c.model_use_robot_state) else None
        s = _lib.RacStep(
            n=n, image=_lib.ptr(image), mask=_lib.ptr(mask), robot=_lib.ptr(r), robot_next=_lib.ptr(r2),
            action=_lib.ptr(action), eps=_lib.ptr(eps), seed=self._seed, noise_ctr=self._noise_ctr,
            sample_mean=int(bool(sample_mean)), use_posterior=int(use_post), next_robot=_lib.ptr(nr),
            eps_post=_lib.ptr(eps_post), force_use_prior=int(bool(force_use_prior)), keep_skip=keep_skip,
            x_pred=_lib.ptr(x_pred), mu_p=_lib.ptr(mu_p), logvar_p=_lib.ptr(logvar_p), mu=_lib.ptr(mu),
            logvar=_lib.ptr(logvar))
        _lib.check(self._lib.rac_forward(self._h, C.byref(s), _lib.stream_ptr()), self._h, "rac_forward")
        self._noise_ctr += 1
        self._eps = self._eps_post = None
        return x_pred, _LazySkips(self, n), mu, logvar, mu_p, logvar_p

    def _write_skips(self, skip, n):
        """Caller-supplied skip tensors (last_frame_skip False): copy into the decoder's concat buffers."""
        for name, shape, off, t in (("cat5", (n, 48, 64, 128), 64, skip[0]), ("cat4", (n, 24, 32, 256), 128, skip[1]),
                                    ("cat3", (n, 12, 16, 512), 256, skip[2])):
            self._buffer_view(name, shape)[..., off:] = t.to(self._device).permute(0, 2, 3, 1).to(torch.bfloat16)
