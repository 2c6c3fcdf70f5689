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
    # sharded == unsharded: the same plan with every candidate on this GPU (process_group=None -> one rac_cem_plan
    # call), same seed: Philox z noise and action noise are keyed on the GLOBAL candidate id, so costs, elites and the
    # refit mean must be bit-equal to the R = 2 plan
    torch.manual_seed(0)
    single = CEMPolicy(cfg, model, horizon=4, opt_iter=4, action_candidates=301, topk=30, init_std=0.03,
                       process_group=None, noise_source="philox")
    single._seed = 1234
    mean1 = single.get_action(start, goal, 0, 0)
    np.testing.assert_array_equal(mean1, out["1"][0])
    np.testing.assert_array_equal(single.last_costs.cpu().numpy(), out["1"][1])
    np.testing.assert_array_equal(single.last_elite_idx.cpu().numpy(), out["1"][2])
    # robot-aware model (configs[4]): precomputed per-candidate states / masks, sharded by pointer offset + time stride
    cfg_ra = so.make_cfg(g_dim=128, z_dim=10, model_use_mask=True, model_use_future_mask=True, model_use_robot_state=True,
                         reconstruction_loss="dontcare_l1", reward_type="dontcare")
    model_ra = SVGConvModel(cfg_ra)
    model_ra.load_state_dict(so.make_state_dict(cfg_ra, 1))
    model_ra.eval()
    gen = torch.Generator(device="cuda").manual_seed(7)
    N, L = 203, 3
    states = torch.rand(L + 1, N, 5, device="cuda", generator=gen)
    masks = (torch.rand(L + 1, N, 1, 48, 64, device="cuda", generator=gen) > 0.8).float()
    goal_ra = DemoGoalState(imgs=goal.imgs, masks=[(rs.rand(1, 48, 64) > 0.8).astype(np.float32)])
    res = []
    for group in (dist.group.WORLD, None):
        pol = CEMPolicy(cfg_ra, model_ra, horizon=L + 1, opt_iter=3, action_candidates=N, topk=20, init_std=0.03,
                        process_group=group, noise_source="philox")
        pol._seed = 99
        pol.precomputed_robot = (states, masks)
        mean = pol.get_action(start, goal_ra, 0, 0)
        res.append((mean, pol.last_costs.cpu().numpy().copy(), pol.last_elite_idx.cpu().numpy().copy()))
    for a, b in zip(res[0], res[1]):
        np.testing.assert_array_equal(a, b)
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, out["1"][0].tolist())
    assert all(g == gathered[0] for g in gathered), "ranks disagree"
    if rank == 0:
        print("peer-memory cost exchange == NCCL all-gather: mean", out["1"][0][0], "(301 candidates, uneven shards)")
        print("sharded plan (R = 2) == unsharded plan (R = 1): costs, elites, mean bit-equal (vanilla 301, robot-aware 203)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
