"""Run under torchrun with 2 GPUs (tests/test_gpu_multi.py): data-parallel SVG training step. Checks that
(a) the all-reduce started underneath the backward pass (grads_ready callback, per-layer ranges) + the remainder gives
    the same averaged gradients as ONE all-reduce of the whole buffer after the backward pass, and both equal the
    mean of the two ranks' local gradients computed by hand;
(b) the Adam update with the 1 / world factor inside the kernel equals torch.optim.Adam on the averaged gradient;
(c) parameters stay bit-identical across ranks after the step;
(d) the fused optimizer step (gradients of the large layers stay packed, are all-reduced in that form and consumed by
    the Adam + re-pack kernel) ends in the same parameters bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svg_oracle as so  # noqa: E402
from robot_aware_control_b200 import SVGConvModel, SVGTrainer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    cfg = so.make_cfg(g_dim=128, z_dim=10, lr=1e-3, beta=1e-2, beta1=0.9, n_future=3, n_past=1)
    sd = so.make_state_dict(cfg, 5)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)  # every rank its own batch
    B, T = 4, 4
    batch = {"images": torch.rand(T, B, 3, 48, 64, device=dev, generator=g),
             "actions": torch.rand(T - 1, B, cfg.action_dim, device=dev, generator=g) * 0.1 - 0.05}
    eps = (torch.randn(T - 1, B, 10, 6, 8, device=dev, generator=g), torch.randn(T - 1, B, 10, 6, 8, device=dev, generator=g))
    out = {}
    import robot_aware_control_b200.trainer as tr
    tr.OVERLAP_MIN_ELEMS = 100_000  # (g128: the gate convolutions have 0.6-1.6 M weights)
    tr.FUSED_MIN_ELEMS = 100_000
    for mode in ("overlap", "plain", "fused", "local"):
        model = SVGConvModel(cfg)
        model.load_state_dict(sd)
        model.train()
        trainer = SVGTrainer(cfg, model, process_group=None if mode == "local" else dist.group.WORLD)
        trainer.overlap_allreduce = mode in ("overlap", "fused")
        trainer.set_noise(*eps)
        trainer.forward_backward(batch, fused_update=mode == "fused")
        n_pending = len(trainer._pending)
        if mode == "local":
            out[mode] = trainer.grads.clone()
            continue
        p0 = trainer.params.clone()
        trainer.optimizer_step()
        out[mode] = (trainer.grads.clone(), trainer.params.clone(), p0, n_pending)
    assert out["overlap"][3] >= 6 and out["plain"][3] == 0, (out["overlap"][3], out["plain"][3])
    both = [torch.empty_like(out["local"]) for _ in range(world)]
    dist.all_gather(both, out["local"])
    summed = both[0] + both[1]
    for mode in ("overlap", "plain"):
        grads, params, p0, _ = out[mode]
        assert torch.equal(grads, summed), mode  # SUM all-reduce of two ranks (the mean is taken inside Adam)
        pt = torch.nn.Parameter(p0.clone())
        opt = torch.optim.Adam([pt], lr=1e-3, betas=(0.9, 0.999))
        pt.grad = summed / world
        opt.step()
        torch.testing.assert_close(params, pt.detach(), rtol=2e-5, atol=2e-7)
        chk = [torch.empty_like(params) for _ in range(world)]
        dist.all_gather(chk, params)
        assert torch.equal(chk[0], chk[1]), mode
    assert torch.equal(out["overlap"][1], out["plain"][1])
    # fused optimizer step (packed gradients all-reduced in place, Adam + re-pack in one pass): the same parameters
    assert out["fused"][3] >= 6 and torch.equal(out["fused"][1], out["plain"][1])
    if rank == 0:
        print("overlapped all-reduce == one all-reduce == mean of the local gradients; replicas identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
