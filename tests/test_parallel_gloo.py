"""World-size-2 `gloo` test (CPU) of the host-side sharding logic used by the multi-GPU planner: contiguous shards,
all-gather of per-candidate fp64 costs in global candidate order (even and uneven splits), and a replicated elite
selection + refit that is bit-identical on every rank. The per-shard costs come from the CPU oracle (tiny model) so
that sharded == unsharded can be asserted without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import svg_oracle as so
    from robot_aware_control_b200 import parallel

    group = dist.group.WORLD
    assert parallel.world_info(group) == (world, rank)
    lo, hi = parallel.shard_range(n_total, rank, world)
    # every rank builds the same plan inputs (replicated), rolls out only its shard
    cfg = so.make_cfg(g_dim=64, z_dim=10, sample_mean=True)
    model = so.SVGOracle(cfg, so.make_state_dict(cfg, 3))
    rs = np.random.RandomState(0)
    start = rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)
    goals = [rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)]
    g = torch.Generator().manual_seed(1)
    L = 2
    act = so.cem_sample(torch.zeros(L, 2), torch.ones(L, 2) * 0.03, torch.randn(n_total, L, 2, generator=g), 0)
    padded = torch.cat([act, torch.zeros(n_total, L, 3)], 2)
    eps = torch.zeros(L, hi - lo, 10, 6, 8)
    local = so.rollout_cost(model, cfg, padded[lo:hi], start, goals, None, None, None, eps)["sum_cost"]
    costs = parallel.all_gather_costs(torch.from_numpy(local), n_total, group)
    assert costs.shape == (n_total,) and costs.dtype == torch.float64
    elite = so.topk_largest(costs.numpy(), max(1, n_total // 3))
    mean, std = so.cem_refit(act, elite)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), costs=costs.numpy(), elite=elite, mean=mean.numpy(),
             std=std.numpy(), lo=lo, hi=hi)
    if rank == 0:
        full = so.rollout_cost(model, cfg, padded, start, goals, None, None, None, torch.zeros(L, n_total, 10, 6, 8))
        np.savez(os.path.join(out_dir, "full.npz"), costs=full["sum_cost"])
    dist.barrier()
    dist.destroy_process_group()


def _run(n_total, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    full = np.load(tmp_path / "full.npz")
    assert int(r0["lo"]) == 0 and int(r0["hi"]) == int(r1["lo"]) and int(r1["hi"]) == n_total
    for k in ("costs", "elite", "mean", "std"):
        np.testing.assert_array_equal(r0[k], r1[k])  # replicated refit: bit-identical on all ranks
    np.testing.assert_allclose(r0["costs"], full["costs"], rtol=1e-6)  # sharded == unsharded


def test_sharded_plan_even_split(tmp_path):
    _run(6, tmp_path)


def test_sharded_plan_uneven_split(tmp_path):
    _run(5, tmp_path)
