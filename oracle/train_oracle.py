"""TEST INFRASTRUCTURE -- CPU oracle of the SVG training step (reference PredictionTrainer._train_step,
src/prediction/trainer.py:326-465 with _recon_loss :149-161 and Adam from _init_models :109-122), restated with
torch autograd on the functional oracle model (oracle/svg_oracle.py, BatchNorm in train mode). Pinned by
tests/golden/train_*.npz, which oracle/make_golden_train.py generates by running the UNMODIFIED reference trainer."""
import numpy as np
import torch

from oracle import svg_oracle as so


class TrainOracle:
    def __init__(self, cfg, state_dict, lr=1e-4, beta1=0.9, beta=1e-4, robot_pixel_weight=0.0):
        self.cfg = cfg
        self.model = so.SVGOracle(cfg, state_dict)
        self.model.bn_training = True
        self.param_keys = [k for k, v in self.model.sd.items()
                           if v.is_floating_point() and "running_" not in k]
        for k in self.param_keys:
            self.model.sd[k] = self.model.sd[k].clone().requires_grad_(True)
        for k in self.model.sd:
            if "running_" in k:
                self.model.sd[k] = self.model.sd[k].clone()
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, 0.999, 1e-8
        self.beta, self.robot_pixel_weight = beta, robot_pixel_weight
        self.m = {k: torch.zeros_like(self.model.sd[k]) for k in self.param_keys}
        self.v = {k: torch.zeros_like(self.model.sd[k]) for k in self.param_keys}
        self.t = 0

    def recon_loss(self, pred, target, mask, batch_weight=None):
        """trainer.py:149-161 (batch_weight reaches the two l1 criteria only, :426-431)."""
        kind = self.cfg.reconstruction_loss
        if kind == "l1":
            return so.l1_criterion(pred, target, batch_weight)
        if kind == "dontcare_l1":
            return so.dontcare_l1_criterion(pred, target, mask, self.robot_pixel_weight, batch_weight)
        if kind == "mse":
            return so.mse_criterion(pred, target)
        if kind == "dontcare_mse":
            return so.dontcare_mse_criterion(pred, target, mask, self.robot_pixel_weight)
        raise NotImplementedError(kind)

    def loss_and_grads(self, batch, eps_prior, eps_post, true_token=None):
        """trainer.py:326-460 up to loss.backward(). batch: images (T,B,3,H,W), masks (T,B,1,H,W), states (T,B,R),
        actions (T-1,B,A), time-first as the reference data loader (robonet_dataset.py:434-451).
        eps_*: (T-1, B, z, 6, 8) injected reparameterisation noise (prior drawn first, then posterior).
        true_token[i]: scheduled-sampling decision for step i >= 2 (None: always ground truth)."""
        cfg, model = self.cfg, self.model
        x, mask, states, ac = batch["images"], batch["masks"], batch["states"], batch["actions"]
        T, B = x.shape[0], x.shape[1]
        for k in self.param_keys:
            model.sd[k].grad = None
        model.init_hidden(B)
        recon = kld = 0
        info = {"recon_loss": 0.0, "kld": 0.0}
        x_pred = None
        skip = None
        dontcare = ("dontcare" in cfg.reconstruction_loss) or cfg.black_robot_input
        for i in range(1, T):
            use_true = i == 1 or true_token is None or bool(true_token[i])
            x_j = x[i - 1] if use_true else x_pred.clone()
            m_j, r_j, a_j = mask[i - 1], states[i - 1], ac[i - 1]
            x_i, m_i, r_i = x[i], mask[i], states[i]
            x_j_black = so.zero_robot_region(m_j, x_j) if dontcare else x_j
            if cfg.last_frame_skip:
                skip = None
            m_in = torch.cat([m_j, m_i], 1) if cfg.model_use_future_mask else m_j
            r_in = (r_j, r_i) if cfg.model_use_future_robot_state else r_j
            out = model.forward_grad(x_j_black, m_in if cfg.model_use_mask else None,
                                     r_in if cfg.model_use_robot_state else None, a_j, eps_prior[i - 1],
                                     next_robot=r_i if cfg.model_use_robot_state else None, eps_post=eps_post[i - 1],
                                     use_posterior=True, skip=skip)
            x4, curr_skip, mu, logvar, mu_p, logvar_p = out
            rgb, m_hat = x4[:, :3], x4[:, 3:4]
            x_pred = (1 - m_hat) * x_j + m_hat * rgb  # blends with the UN-blacked x_j (trainer.py:406-407)
            if i <= 1:  # n_past == 1
                skip = curr_skip
            bw = None
            if getattr(cfg, "load_movement_info", False):  # trainer.py:426-429
                info_m = batch["high_movement"]
                bw = (cfg.movement_weight * info_m).float()
                bw[~info_m] = 1.0
            vl = self.recon_loss(x_pred, x_i, m_i, bw)
            recon = recon + vl
            kl = so.kl_criterion(mu, logvar, mu_p, logvar_p, B)
            kld = kld + kl
            info["recon_loss"] += float(vl.detach())
            info["kld"] += float(kl.detach())
        loss = recon + kld * self.beta
        loss.backward()
        info["loss"] = float(loss.detach())
        return info, {k: model.sd[k].grad.detach().clone() for k in self.param_keys}

    @torch.no_grad()
    def adam_step(self, grads):
        """torch.optim.Adam (no weight decay, no amsgrad) as constructed at trainer.py:109-122."""
        self.t += 1
        b1, b2 = self.beta1, self.beta2
        for k in self.param_keys:
            g = grads[k]
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
            denom = (self.v[k].sqrt() / np.sqrt(bc2)).add_(self.eps)
            self.model.sd[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)

    def train_step(self, batch, eps_prior, eps_post, true_token=None):
        info, grads = self.loss_and_grads(batch, eps_prior, eps_post, true_token)
        self.adam_step(grads)
        return info, grads
