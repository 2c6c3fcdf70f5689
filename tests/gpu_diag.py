"""TEST INFRASTRUCTURE -- layer-by-layer diagnosis on a B200: runs one SVG step through libracb200.so with the SIMT
cross-check kernels and with the tcgen05 kernels and prints, per internal buffer, the max abs difference to the CPU
oracle (and tc vs simt). Usage on the GPU box:  python -m tests.gpu_diag [g_dim] [batch]"""
import sys

import torch

from oracle import svg_oracle as so
from robot_aware_control_b200 import SVGConvModel

NHWC = lambda t: t.permute(0, 2, 3, 1).contiguous()


def run(impl, cfg, sd, d, B):
    torch.manual_seed(0)
    m = SVGConvModel(cfg, conv_impl=impl)
    m.load_state_dict(sd)
    m.eval()
    m.init_hidden(B)
    g = cfg.g_dim
    bufs = {}
    outs = []
    for t in range(2):
        m.set_noise(eps=d["eps"][t])
        out = m.forward(d["image"][t], None, None, None, d["action"][t])
        torch.cuda.synchronize()
        outs.append([o.float().cpu() if isinstance(o, torch.Tensor) else None for o in (out[0], out[4], out[5])])
        if t == 0:
            for name, shape in (("a1", (B, 48, 64, 64)), ("cat5", (B, 48, 64, 128)), ("a2", (B, 24, 32, 128)),
                                ("cat4", (B, 24, 32, 256)), ("cat3", (B, 12, 16, 512)), ("h4", (B, 6, 8, g)),
                                ("prior_in", (B, 6, 8, g)), ("prior.h0.1", (B, 6, 8, g)), ("prior.h1.1", (B, 6, 8, g)),
                                ("z", (B, 6, 8, 64)), ("frame_in", (B, 6, 8, g)), ("fp.h0.1", (B, 6, 8, g)),
                                ("fp.h1.1", (B, 6, 8, g)), ("d2a", (B, 6, 8, 512)), ("d2b", (B, 6, 8, 512)),
                                ("d5", (B, 48, 64, 64))):
                bufs[name] = m._buffer_view(name, shape).float().cpu()
            bufs["prior.c0"] = m._buffer_view("prior.c0", (B, 6, 8, g), torch.float32).cpu()
    return bufs, outs


def main():
    g_dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    impls = sys.argv[3].split(",") if len(sys.argv) > 3 else ["simt", "tc"]
    cfg = so.make_cfg(g_dim=g_dim, z_dim=10)
    sd = so.make_state_dict(cfg, 11)
    gen = torch.Generator().manual_seed(21)
    d = {"image": torch.rand(2, B, 3, 48, 64, generator=gen),
         "action": (torch.rand(2, B, cfg.action_dim, generator=gen) - 0.5) * 0.1,
         "eps": torch.randn(2, B, cfg.z_dim, 6, 8, generator=gen)}
    oracle = so.SVGOracle(cfg, sd)
    oracle.init_hidden(B)
    oracle.trace = {}
    ref_out = []
    tr0 = None
    for t in range(2):
        o = oracle.forward(d["image"][t], None, None, d["action"][t], d["eps"][t])
        ref_out.append((o[0], o[4], o[5]))
        if t == 0:
            tr0 = dict(oracle.trace)
    tr = tr0
    zpad = torch.zeros(B, 6, 8, 64)
    zpad[..., :cfg.z_dim] = NHWC(tr["z"])
    ref = {
        "a1": NHWC(tr["a1"]), "cat5[64:]": NHWC(tr["h1"]), "a2": NHWC(tr["a2"]), "cat4[128:]": NHWC(tr["h2"]),
        "cat3[256:]": NHWC(tr["h3"]), "h4": NHWC(tr["h4"]), "prior_in": NHWC(tr["prior_in"]),
        "prior.h0.1": NHWC(tr["prior.h0"]), "prior.c0": NHWC(tr["prior.c0"]), "prior.h1.1": NHWC(tr["prior.h1"]),
        "z": zpad, "frame_in": NHWC(tr["frame_in"]), "fp.h0.1": NHWC(tr["frame_predictor.h0"]),
        "fp.h1.1": NHWC(tr["frame_predictor.h1"]), "d2a": NHWC(tr["d2.0"]), "d2b": NHWC(tr["d2.1"]),
        "d5": NHWC(tr["d5"]),
    }
    results = {}
    for impl in impls:
        try:
            results[impl] = run(1 if impl == "simt" else 0, cfg, sd, d, B)
        except Exception as e:  # keep going: the other implementation still tells us something
            print(f"[{impl}] FAILED: {type(e).__name__}: {e}")
    slices = {"cat5[64:]": ("cat5", 64), "cat4[128:]": ("cat4", 128), "cat3[256:]": ("cat3", 256)}
    print(f"{'buffer':14s} " + " ".join(f"{i + ' vs oracle':>16s}" for i in results) + ("   tc vs simt" if len(results) == 2 else ""))
    for name, r in ref.items():
        row = []
        got = {}
        for impl, (bufs, _) in results.items():
            if name in slices:
                b, off = slices[name]
                x = bufs[b][..., off:]
            else:
                x = bufs[name]
            got[impl] = x
            row.append(f"{(x - r).abs().max().item():16.5f}")
        extra = ""
        if len(got) == 2:
            extra = f"   {(got['tc'] - got['simt']).abs().max().item():.5f}"
        print(f"{name:14s} " + " ".join(row) + extra + f"   (ref absmax {r.abs().max().item():.3f})")
    for t in range(2):
        for i, nm in enumerate(("x_pred", "mu_p", "logvar_p")):
            row = [f"{(outs[t][i] - ref_out[t][i]).abs().max().item():16.5f}" for _, (_, outs) in results.items()]
            print(f"step{t} {nm:8s} " + " ".join(row))


if __name__ == "__main__":
    main()
