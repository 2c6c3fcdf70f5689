"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest): the fused peer-memory cost exchange
(cost kernel stores into every rank's gathered vector + flag barrier) gives bit-identical plans to the NCCL all-gather.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_peer_gather_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svg_oracle as so  # noqa: E402
from robot_aware_control_b200 import CEMPolicy, DemoGoalState, State, SVGConvModel  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = so.make_cfg(g_dim=128, z_dim=10)
    model = SVGConvModel(cfg)
    model.load_state_dict(so.make_state_dict(cfg, 0))
    model.eval()
    rs = np.random.RandomState(0)
    start = State(img=rs.randint(0, 256, (48, 64, 3)).astype(np.uint8))
    goal = DemoGoalState(imgs=[rs.randint(0, 256, (48, 64, 3)).astype(np.uint8)], masks=[np.zeros((1, 48, 64), np.float32)])
    out = {}
    for mode in ("0", "1"):
        os.environ["RAC_PEER_GATHER"] = mode
        torch.manual_seed(0)
        policy = CEMPolicy(cfg, model, horizon=4, opt_iter=4, action_candidates=301, topk=30, init_std=0.03,
                           process_group=dist.group.WORLD, noise_source="philox")
        policy._seed = 1234
        mean = policy.get_action(start, goal, 0, 0)
        out[mode] = (mean, policy.last_costs.cpu().numpy().copy(), policy.last_elite_idx.cpu().numpy().copy())
        if mode == "1":
            assert getattr(policy, "_peer_exchange", None) is not None, "peer exchange was not used"
    np.testing.assert_array_equal(out["0"][0], out["1"][0])
    np.testing.assert_array_equal(out["0"][1], out["1"][1])
    np.testing.assert_array_equal(out["0"][2], out["1"][2])
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, out["1"][0].tolist())
    assert all(g == gathered[0] for g in gathered), "ranks disagree"
    if rank == 0:
        print("peer-memory cost exchange == NCCL all-gather: mean", out["1"][0][0], "(301 candidates, uneven shards)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
