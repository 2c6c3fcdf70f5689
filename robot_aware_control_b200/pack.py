"""Weight packing: reference `state_dict` -> the K-major bf16 GEMM operands the sm_100a kernels consume.

Per convolution the packed weight is Wp[n, tap, c] (n = output column, tap = kh * ks + kw, c = concatenated padded
input channels), i.e. row-major [n_packed, ks*ks*ctot] = the "B" operand of the implicit GEMM in csrc/conv_tc.cu.
Identities used:
  * eval BatchNorm fold into the bias-free vgg_layer conv (reference vgg_64.py:8-18)
  * ConvTranspose2d(64, 4, 3, 1, 1) == conv2d with transposed + flipped weight (vgg_64.py:219)
  * tiled action / robot-state channels (dynamics.py:591-603) live in one zero-padded 64-channel block
  * LSTM gate columns interleaved (channel, gate) so that one accumulator row chunk holds in/remember/out/cell of
    the same channel (lstm.py:135); mu/logvar columns interleaved (z channel, {mu, logvar}) (lstm.py:283-285)
"""
from collections import OrderedDict

import torch

# layer ids: must match the enum in include/racb200.h
LAYER_IDS = [
    "ENC_C1_0", "ENC_C1_1", "ENC_C2_0", "ENC_C2_1", "ENC_C3_0", "ENC_C3_1", "ENC_C3_2", "ENC_C4_0", "ENC_C4_1",
    "ENC_C4_2", "PRIOR_IN", "PRIOR_LSTM0", "PRIOR_LSTM1", "PRIOR_GAUSS", "FP_IN", "FP_LSTM0", "FP_LSTM1",
    "DEC_UPC2_0", "DEC_UPC2_1", "DEC_UPC2_2", "DEC_UPC3_0", "DEC_UPC3_1", "DEC_UPC3_2", "DEC_UPC4_0", "DEC_UPC4_1",
    "DEC_UPC5_0", "DEC_UPC5_1", "POST_IN", "POST_LSTM0", "POST_LSTM1", "POST_GAUSS",
]
# cfg.lstm_group_norm only (RAC_L_*_HH): hh_gates convolutions of the NormConvLSTMCells (lstm.py:168-171)
GN_LAYER_IDS = ["PRIOR_LSTM0_HH", "PRIOR_LSTM1_HH", "FP_LSTM0_HH", "FP_LSTM1_HH", "POST_LSTM0_HH", "POST_LSTM1_HH"]
LAYER_INDEX = {n: i for i, n in enumerate(LAYER_IDS + GN_LAYER_IDS)}

_VGG = OrderedDict([
    ("ENC_C1_1", "encoder.c1.1"), ("ENC_C2_0", "encoder.c2.0"), ("ENC_C2_1", "encoder.c2.1"),
    ("ENC_C3_0", "encoder.c3.0"), ("ENC_C3_1", "encoder.c3.1"), ("ENC_C3_2", "encoder.c3.2"),
    ("ENC_C4_0", "encoder.c4.0"), ("ENC_C4_1", "encoder.c4.1"), ("ENC_C4_2", "encoder.c4.2"),
    ("DEC_UPC2_0", "decoder.upc2.0"), ("DEC_UPC2_1", "decoder.upc2.1"), ("DEC_UPC2_2", "decoder.upc2.2"),
    ("DEC_UPC3_0", "decoder.upc3.0"), ("DEC_UPC3_1", "decoder.upc3.1"), ("DEC_UPC3_2", "decoder.upc3.2"),
    ("DEC_UPC4_0", "decoder.upc4.0"), ("DEC_UPC4_1", "decoder.upc4.1"), ("DEC_UPC5_0", "decoder.upc5.0"),
])


def _round_up(a, b):
    return (a + b - 1) // b * b


def fold_bn(sd, prefix, eps=1e-5):
    """vgg_layer in eval mode == conv(W * s) + (beta - mean * s), s = gamma / sqrt(var + eps)."""
    w = sd[f"{prefix}.main.0.weight"].float()
    s = sd[f"{prefix}.main.1.weight"].float() / torch.sqrt(sd[f"{prefix}.main.1.running_var"].float() + eps)
    b = sd[f"{prefix}.main.1.bias"].float() - sd[f"{prefix}.main.1.running_mean"].float() * s
    return w * s[:, None, None, None], b


def pack_gemm(w, bias, splits, block_n, col_src=None, n_valid=None):
    """w (cout, cin, k, k) fp32, bias (cout) -> (Wp bf16 [n_packed, k*k*ctot], bias fp32 [n_packed]).
    splits: list of (real_channels, padded_channels) per concatenated input, in the reference's cat order.
    col_src: packed column -> source output channel (or -1 for a zero column)."""
    cout, cin, k, _ = w.shape
    assert sum(r for r, _ in splits) == cin, (cin, splits)
    ctot = sum(p for _, p in splits)
    if col_src is None:
        col_src = list(range(cout))
    n_packed = _round_up(len(col_src), block_n)
    wp = torch.zeros(n_packed, k * k, ctot, dtype=torch.float32)
    bp = torch.zeros(n_packed, dtype=torch.float32)
    src = torch.tensor([c for c in col_src if c >= 0], dtype=torch.long)
    dst = torch.tensor([i for i, c in enumerate(col_src) if c >= 0], dtype=torch.long)
    wt = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)  # (cout, tap, cin)
    ci = co = 0
    for real, padded in splits:
        wp[dst, :, co:co + real] = wt[src, :, ci:ci + real]
        ci += real
        co += padded
    bp[dst] = bias.float()[src]
    return wp.reshape(n_packed, k * k * ctot).to(torch.bfloat16).contiguous(), bp.contiguous()


def pack_state_dict(sd, cfg):
    """Returns OrderedDict layer name -> (weight tensor, bias tensor) in the order of LAYER_IDS."""
    g, z, a, r = cfg.g_dim, cfg.z_dim, cfg.action_dim, cfg.robot_dim
    use_r = bool(cfg.model_use_robot_state)
    use_r2 = use_r and bool(cfg.model_use_future_robot_state)
    naux = a + (r if use_r else 0) + (r if use_r2 else 0)
    out = OrderedDict()

    # encoder.c1.0: fp32 [9 * cin, 64], tap-major (SIMT first layer, csrc/misc_kernels.cu)
    w, b = fold_bn(sd, "encoder.c1.0")
    cin = w.shape[1]
    out["ENC_C1_0"] = (w.permute(2, 3, 1, 0).reshape(9 * cin, 64).contiguous().float(), b.contiguous())

    for name, prefix in _VGG.items():
        w, b = fold_bn(sd, prefix)
        cout, cin = w.shape[0], w.shape[1]
        block_n = 64 if cout == 64 else 128
        out[name] = pack_gemm(w, b, [(cin, cin)], block_n)

    def conv(prefix):
        return sd[f"{prefix}.weight"].float(), sd[f"{prefix}.bias"].float()

    w, b = conv("prior_input_conv")
    out["PRIOR_IN"] = pack_gemm(w, b, [(naux, 64), (g, g)], 128)
    w, b = conv("frame_pred_input_conv")
    out["FP_IN"] = pack_gemm(w, b, [(naux, 64), (g, g), (z, 64)], 128)
    w, b = conv("posterior_input_conv")
    out["POST_IN"] = pack_gemm(w, b, ([(r, 64)] if use_r else []) + [(g, g)], 128)

    gate_cols = [gate * g + ch for ch in range(g) for gate in range(4)]  # packed col = ch * 4 + gate
    group_norm = bool(getattr(cfg, "lstm_group_norm", False))
    for tag, prefix in (("PRIOR", "prior"), ("POST", "posterior"), ("FP", "frame_predictor")):
        for layer in (0, 1):
            if group_norm:  # NormConvLSTMCell: separate ih / hh convolutions (lstm.py:163-171)
                w, b = conv(f"{prefix}.lstm.{layer}.ih_gates.0")
                out[f"{tag}_LSTM{layer}"] = pack_gemm(w, b, [(g, g)], 128, col_src=gate_cols)
                w, b = conv(f"{prefix}.lstm.{layer}.hh_gates.0")
                out[f"{tag}_LSTM{layer}_HH"] = pack_gemm(w, b, [(g, g)], 128, col_src=gate_cols)
            else:
                w, b = conv(f"{prefix}.lstm.{layer}.gates")
                out[f"{tag}_LSTM{layer}"] = pack_gemm(w, b, [(g, g), (g, g)], 128, col_src=gate_cols)
    for tag, prefix in (("PRIOR", "prior"), ("POST", "posterior")):
        wm, bm = conv(f"{prefix}.mu_net")
        wl, bl = conv(f"{prefix}.logvar_net")
        w = torch.cat([wm, wl], 0)
        b = torch.cat([bm, bl], 0)
        cols = []
        for zc in range(64):
            cols += [zc, z + zc] if zc < z else [-1, -1]
        out[f"{tag}_GAUSS"] = pack_gemm(w, b, [(g, g)], 128, col_src=cols)

    wt = sd["decoder.upc5.1.weight"].float()  # ConvTranspose2d weight (in=64, out=4, 3, 3)
    wc = wt.transpose(0, 1).flip(2, 3).contiguous()
    out["DEC_UPC5_1"] = pack_gemm(wc, sd["decoder.upc5.1.bias"].float(), [(64, 64)], 16)

    return OrderedDict((n, out[n]) for n in LAYER_IDS + (GN_LAYER_IDS if group_norm else []))


def pack_lstm_norm(sd, cfg):
    """cfg.lstm_group_norm: OrderedDict cell layer name -> fp32 [18 g] GroupNorm affine parameters of that
    NormConvLSTMCell for rac_load_lstm_norm: ih / hh gamma and beta in packed gate-column order, then c_norm."""
    g = cfg.g_dim
    cols = torch.tensor([gate * g + ch for ch in range(g) for gate in range(4)], dtype=torch.long)
    out = OrderedDict()
    for tag, prefix in (("PRIOR", "prior"), ("POST", "posterior"), ("FP", "frame_predictor")):
        for layer in (0, 1):
            p = f"{prefix}.lstm.{layer}"
            parts = [sd[f"{p}.{gk}.1.{wb}"].float()[cols] for gk in ("ih_gates", "hh_gates") for wb in ("weight", "bias")]
            parts += [sd[f"{p}.c_norm.weight"].float(), sd[f"{p}.c_norm.bias"].float()]
            out[f"{tag}_LSTM{layer}"] = torch.cat(parts).contiguous()
    return out
