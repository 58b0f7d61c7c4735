"""Multi-GPU plumbing: environments are independent, so a job of ``total_envs`` is split into
contiguous ranges of *global env id*, one per rank (one process per GPU), with no data-path
collective.  The only exchange is the all-reduce of the per-rank counter sums at rollout /
episode boundaries (global blocking rate, mean reward) -- what SB3's Monitor logs per episode
(SURVEY.md section 8e)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int):
    """[first, first+count) of global env ids owned by ``rank``; remainders go to the low ranks."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def global_statistics(local_sums: torch.Tensor, group=None) -> dict:
    """All-reduce (sum) of the int64 [9] vector from ``OpticalVecEnv.reduce_counters`` and the
    job-wide blocking rates derived from it.  Works on NCCL (CUDA tensors) and gloo (CPU tensors)."""
    sums = local_sums.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    s = [int(x) for x in sums.tolist()]
    out = {
        "services_processed": s[0], "services_accepted": s[1],
        "episode_services_processed": s[2], "episode_services_accepted": s[3],
        "bit_rate_requested": s[4], "bit_rate_provisioned": s[5],
        "episode_bit_rate_requested": s[6], "episode_bit_rate_provisioned": s[7],
        "envs_with_errors": s[8],
        "service_blocking_rate": (s[0] - s[1]) / s[0] if s[0] else 0.0,
        "bit_rate_blocking_rate": (s[4] - s[5]) / s[4] if s[4] else 0.0,
    }
    return out
