"""Candidate sharding for multi-GPU CEM (new: the reference is single-process).

One process per GPU (`torch.distributed`, NCCL over NVLink). Rank r owns the contiguous candidate range
shard_range(N, r, R); weights, start/goal images and the sampling distribution are replicated; sampling and z noise are
counter-based on the GLOBAL candidate id, so results do not depend on R. The only exchange per CEM iteration is an
all-gather of the per-candidate fp64 costs (N/R values per rank); the top-k and refit are then replicated and
bit-identical on every rank because their input is identical.
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    """(world_size, rank) of `group`; (1, 0) when torch.distributed is not initialised and no group is given."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    if group is None:
        return 1, 0  # a default group exists but the caller did not ask for sharding
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(n, rank, world):
    """Contiguous, balanced split: the first n % world ranks get one extra candidate."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_costs(local_costs, n_total, group=None):
    """Concatenation over ranks of the local cost vectors (fp64), in rank order == global candidate order."""
    world, rank = world_info(group)
    if world == 1:
        return local_costs
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local_costs.numel() != hi - lo:
        raise ValueError(f"rank {rank} holds {local_costs.numel()} costs, its shard has {hi - lo}")
    if all(b - a == width for a, b in sizes):
        out = torch.empty(n_total, dtype=local_costs.dtype, device=local_costs.device)
        dist.all_gather_into_tensor(out, local_costs.contiguous(), group=group)
        return out
    # uneven split: collectives need equal sizes, so pad every shard to the widest one and cut the padding out
    padded = torch.zeros(width, dtype=local_costs.dtype, device=local_costs.device)
    padded[: hi - lo] = local_costs
    out = torch.empty(world * width, dtype=local_costs.dtype, device=local_costs.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width: r * width + (b - a)] for r, (a, b) in enumerate(sizes)])
