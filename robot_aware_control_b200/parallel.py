"""Candidate sharding for multi-GPU CEM (new: the reference is single-process).

One process per GPU (`torch.distributed`, NCCL over NVLink). Rank r owns the contiguous candidate range
shard_range(N, r, R); weights, start/goal images and the sampling distribution are replicated; sampling and z noise are
counter-based on the GLOBAL candidate id, so results do not depend on R. The only exchange per CEM iteration is an
all-gather of the per-candidate fp64 costs (N/R values per rank); the top-k and refit are then replicated and
bit-identical on every rank because their input is identical.
"""
import os

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


class PeerCostExchange:
    """The cost all-gather without a collective call: two gathered-cost vectors (N fp64, alternating per CEM
    iteration) in torch symmetric memory, mapped into every rank of the node. The cost kernel of the last rollout step
    stores each finished cost into all ranks' vectors over NVLink peer memory (rac_rollout.peer_cost_bufs) and
    rac_peer_barrier exchanges one flag per rank. Two buffers suffice: a rank can run at most one iteration ahead of
    the slowest one (it needs everybody's flag of iteration i + 1 before it can start pushing iteration i + 2).

    Used on NCCL groups of one node; RAC_PEER_GATHER=0, a non-NCCL backend or a failing rendezvous fall back to
    all_gather_costs()."""

    @staticmethod
    def get(owner, n_total, group, device):
        if os.environ.get("RAC_PEER_GATHER", "1") == "0" or group is None or dist.get_backend(group) != "nccl":
            return None
        cached = getattr(owner, "_peer_exchange", None)
        if cached is not None and cached.n_total == n_total and cached.group is group:
            return cached
        if getattr(owner, "_peer_exchange_failed", False):
            return None
        try:
            ex = PeerCostExchange(n_total, group, device)
        except Exception as e:  # no symmetric memory on this system: keep the NCCL path
            owner._peer_exchange_failed = True
            if os.environ.get("RAC_PEER_GATHER") == "1":
                raise
            import warnings

            warnings.warn(f"peer-memory cost exchange unavailable ({e}); using NCCL all_gather")
            return None
        owner._peer_exchange = ex
        return ex

    def __init__(self, n_total, group, device):
        import torch.distributed._symmetric_memory as symm_mem

        self.n_total, self.group = n_total, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty(n_total, dtype=torch.float64, device=device)
            self.handles.append(symm_mem.rendezvous(t, group))
            self.bufs.append(t)
        self.seq = 0
        self.cur = 0
        # the flags live in the upper half of the signal pad (torch's own barriers use the low words)
        self.slot_base = (int(self.handles[0].signal_pad_size) // 4) // 2
        if self.slot_base + self.world > int(self.handles[0].signal_pad_size) // 4:
            raise RuntimeError("signal pad too small")

    def target(self, offset):
        """Arguments for rac_rollout of the next iteration (switches to the other buffer)."""
        self.cur ^= 1
        return (self.handles[self.cur].buffer_ptrs_dev, self.world, offset)

    def finish(self, lib):
        """Flag barrier on the caller's stream; returns this rank's gathered cost vector."""
        from . import _lib

        self.seq += 1
        h = self.handles[0]  # one set of signal pads serves both buffers
        _lib.check(lib.rac_peer_barrier(int(h.signal_pad_ptrs_dev), self.slot_base, self.rank, self.world, self.seq & 0xFFFFFFFF,
                                        _lib.stream_ptr()), None, "rac_peer_barrier")
        return self.bufs[self.cur]
