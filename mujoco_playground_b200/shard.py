"""Multi-GPU sharding of the environment batch: environments never interact, so rank r of R owns a
contiguous slice and there is NO collective on the step path.  The only exchange is the reduction of rollout
statistics (episode count, return sum ...), one small all-reduce per rollout.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the global environment ids owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total_envs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank Philox key: distinct streams per shard, reproducible for a given (seed, world layout)."""
    return (int(seed) * 0x9E3779B97F4A7C15 + rank * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF


_KEYS = ("episodes", "successes", "env_steps", "collisions", "unsupported", "solver_iters", "obstacle_steps", "contacts_sum",
         "bad_state", "return_sum", "length_sum")


def reduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """Sum the per-rank ackb_stats dictionaries over the process group (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    t = torch.tensor([float(stats.get(k, 0)) for k in _KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = {k: float(v) for k, v in zip(_KEYS, t.tolist())}
    for k in _KEYS[:9]:
        out[k] = int(round(out[k]))
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing convention of bench.py: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
