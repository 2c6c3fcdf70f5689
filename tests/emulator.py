"""TEST INFRASTRUCTURE -- CPU emulation of the DATA FLOW of csrc/rac_api.cu::run_step on the packed weights
(robot_aware_control_b200/pack.py): NHWC buffers, channel-offset concat buffers, 64-channel aux block, interleaved
LSTM / gaussian columns, flipped ConvTranspose. It checks the packer and the op graph against the oracle without a
GPU, and (round_bf16=True) predicts the bf16 rounding error the CUDA path should show."""
import torch
import torch.nn.functional as F

from robot_aware_control_b200 import pack


def _q(x, on):
    return x.to(torch.bfloat16).float() if on else x


class PackedEmulator:
    def __init__(self, cfg, state_dict, round_bf16=False):
        self.cfg = cfg
        self.rb = round_bf16
        self.layers = pack.pack_state_dict(state_dict, cfg)
        self.norm = pack.pack_lstm_norm(state_dict, cfg) if getattr(cfg, "lstm_group_norm", False) else None
        self.state = None

    def _gemm(self, name, srcs, ks=3):
        wp, bias = self.layers[name]
        x = torch.cat([_q(s, self.rb) for s in srcs], -1)
        n, ctot = wp.shape[0], x.shape[-1]
        assert wp.shape[1] == ks * ks * ctot, (name, wp.shape, ctot)
        w = wp.float().reshape(n, ks, ks, ctot).permute(0, 3, 1, 2)
        y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, 1, ks // 2)
        return y.permute(0, 2, 3, 1)

    def _act(self, name, srcs, cout, lrelu=True):
        y = self._gemm(name, srcs)[..., :cout]
        if lrelu:
            y = F.leaky_relu(y, 0.2)
        return _q(y, self.rb)

    @staticmethod
    def _up(x):
        return x.repeat_interleave(2, 1).repeat_interleave(2, 2)

    @staticmethod
    def _pool(x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)

    def init_hidden(self, B):
        g = self.cfg.g_dim
        z = lambda: torch.zeros(B, 6, 8, g)
        self.state = {k: [[z(), z()], [z(), z()]] for k in ("PRIOR", "POST", "FP")}

    @staticmethod
    def _gn_gates(raw, gamma, beta, g):
        """csrc/norm_lstm.cu + EPI_GATES: raw [B,6,8,4g] in packed (channel, gate) column order; GroupNorm(16, 4g) groups
        of the reference = (gate, quarter of the channels); gamma / beta in packed column order."""
        B = raw.shape[0]
        r = raw.reshape(B, 48, g, 4)                       # (pos, ch, gate)
        q = g // 4
        r5 = r.reshape(B, 48, 4, q, 4)                      # (pos, quarter, ch in quarter, gate)
        mean = r5.mean(dim=(1, 3), keepdim=True)
        var = r5.var(dim=(1, 3), unbiased=False, keepdim=True)
        n = ((r5 - mean) / torch.sqrt(var + 1e-5)).reshape(B, 48, g, 4)
        return (n * gamma.reshape(1, 1, g, 4) + beta.reshape(1, 1, g, 4)).reshape(B, 6, 8, g, 4)

    def _lstm_gn(self, tag, x):
        """NormConvLSTMCell data flow (lstm.py:177-198) on the packed ih / hh operands and norm vector."""
        g = self.cfg.g_dim
        for layer, ks in ((0, 5), (1, 3)):
            h_prev, c_prev = self.state[tag][layer]
            v = self.norm[f"{tag}_LSTM{layer}"]
            ih = self._gemm(f"{tag}_LSTM{layer}", [x], ks)[..., :4 * g]
            hh = self._gemm(f"{tag}_LSTM{layer}_HH", [h_prev], ks)[..., :4 * g]
            acc = (self._gn_gates(ih, v[0:4 * g], v[4 * g:8 * g], g) +
                   self._gn_gates(hh, v[8 * g:12 * g], v[12 * g:16 * g], g))
            i, f, o, gg = (acc[..., k] for k in range(4))
            c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(gg)
            c16 = c.reshape(c.shape[0], 48, 16, g // 16)
            c16 = (c16 - c16.mean(dim=(1, 3), keepdim=True)) / torch.sqrt(c16.var(dim=(1, 3), unbiased=False, keepdim=True) + 1e-5)
            c = c16.reshape(c.shape) * v[16 * g:17 * g] + v[17 * g:18 * g]
            h = _q(torch.sigmoid(o) * torch.tanh(c), self.rb)
            self.state[tag][layer] = [h, c]
            x = h
        return x

    def _lstm(self, tag, x):
        if self.norm is not None:
            return self._lstm_gn(tag, x)
        g = self.cfg.g_dim
        for layer, ks in ((0, 5), (1, 3)):
            h_prev, c_prev = self.state[tag][layer]
            acc = self._gemm(f"{tag}_LSTM{layer}", [x, h_prev], ks)[..., :4 * g]
            acc = acc.reshape(*acc.shape[:3], g, 4)
            i, f, o, gg = (acc[..., k] for k in range(4))
            c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(gg)
            h = _q(torch.sigmoid(o) * torch.tanh(c), self.rb)
            self.state[tag][layer] = [h, c]
            x = h
        return x

    def _gauss(self, tag, h, eps, sample_mean=False):
        zd = self.cfg.z_dim
        acc = self._gemm(f"{tag}_GAUSS", [h]).reshape(h.shape[0], 6, 8, 64, 2)
        mu, lv = acc[..., :zd, 0], acc[..., :zd, 1]
        z = mu if sample_mean else eps.permute(0, 2, 3, 1) * torch.exp(0.5 * lv) + mu
        zp = torch.zeros(h.shape[0], 6, 8, 64)
        zp[..., :zd] = z
        return _q(zp, self.rb), mu.permute(0, 3, 1, 2), lv.permute(0, 3, 1, 2)

    @torch.no_grad()
    def forward(self, image, mask, robot, action, eps, next_robot=None, eps_post=None, use_posterior=False,
                force_use_prior=False, sample_mean=False):
        cfg = self.cfg
        B = image.shape[0]
        # first layer (fp32 SIMT kernel): [9*cin, 64] tap-major
        w0, b0 = self.layers["ENC_C1_0"]
        x = torch.cat([image, mask], 1) if cfg.model_use_mask else image
        cin = x.shape[1]
        wc = w0.reshape(3, 3, cin, 64).permute(3, 2, 0, 1)
        a1 = _q(F.leaky_relu(F.conv2d(x, wc, b0, 1, 1), 0.2).permute(0, 2, 3, 1), self.rb)
        cat5 = torch.zeros(B, 48, 64, 128)
        cat4 = torch.zeros(B, 24, 32, 256)
        cat3 = torch.zeros(B, 12, 16, 512)
        cat5[..., 64:] = self._act("ENC_C1_1", [a1], 64)
        a2 = self._act("ENC_C2_0", [self._pool(cat5[..., 64:])], 128)
        cat4[..., 128:] = self._act("ENC_C2_1", [a2], 128)
        a3 = self._act("ENC_C3_0", [self._pool(cat4[..., 128:])], 256)
        a3 = self._act("ENC_C3_1", [a3], 256)
        cat3[..., 256:] = self._act("ENC_C3_2", [a3], 256)
        a4 = self._act("ENC_C4_0", [self._pool(cat3[..., 256:])], 512)
        a4 = self._act("ENC_C4_1", [a4], 512)
        h4 = self._act("ENC_C4_2", [a4], cfg.g_dim)
        aux = torch.zeros(B, 6, 8, 64)
        vals = [action]
        if cfg.model_use_robot_state:
            if cfg.model_use_future_robot_state:
                vals += [robot[0], robot[1]]
            else:
                vals += [robot]
        v = torch.cat(vals, 1)
        aux[..., :v.shape[1]] = v[:, None, None, :]
        pin = self._act("PRIOR_IN", [aux, h4], cfg.g_dim, lrelu=False)
        z, mu_p, lv_p = self._gauss("PRIOR", self._lstm("PRIOR", pin), eps, sample_mean)
        mu = lv = None
        if use_posterior:
            srcs = [h4]
            if cfg.model_use_robot_state:
                auxp = torch.zeros(B, 6, 8, 64)
                auxp[..., :next_robot.shape[1]] = next_robot[:, None, None, :]
                srcs = [auxp, h4]
            postin = self._act("POST_IN", srcs, cfg.g_dim, lrelu=False)
            z_t, mu, lv = self._gauss("POST", self._lstm("POST", postin), eps_post)
            if not force_use_prior:
                z = z_t
        fin = self._act("FP_IN", [aux, h4, z], cfg.g_dim, lrelu=False)
        hp = self._lstm("FP", fin)
        d = self._act("DEC_UPC2_0", [hp], 512)
        d = self._act("DEC_UPC2_1", [d], 512)
        cat3[..., :256] = self._up(self._act("DEC_UPC2_2", [d], 256))
        d = self._act("DEC_UPC3_0", [cat3], 256)
        d = self._act("DEC_UPC3_1", [d], 256)
        cat4[..., :128] = self._up(self._act("DEC_UPC3_2", [d], 128))
        d = self._act("DEC_UPC4_0", [cat4], 128)
        cat5[..., :64] = self._up(self._act("DEC_UPC4_1", [d], 64))
        d5 = self._act("DEC_UPC5_0", [cat5], 64)
        x_pred = torch.sigmoid(self._gemm("DEC_UPC5_1", [d5])[..., :4]).permute(0, 3, 1, 2)
        return x_pred, None, mu, lv, mu_p, lv_p
