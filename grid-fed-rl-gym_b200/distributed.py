"""Multi-GPU plumbing of the batched step path.

Instances shard trivially: rank r owns a contiguous range of global instance ids and keys its
Philox streams with ``seed + global id``, so results do not depend on the number of ranks and the
step path carries **no collective**.  The only communication is the optional episode-statistics
reduction (one all-reduce of a short fp64 vector: NCCL over NVLink on GPUs, gloo in the CPU tests)
and the max-over-ranks of a timed region in ``bench.py``.
"""

from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("reward_sum", "episode_reward_sum", "converged", "iterations_sum", "violation_steps",
             "done", "errors", "num_envs")


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous instance range [start, stop) of ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank must be in [0, world_size)")
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_reduce_stats(local: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank statistics vectors (fp64[len(STAT_KEYS)]), in place."""
    if local.dtype != torch.float64 or local.numel() != len(STAT_KEYS):
        raise ValueError(f"expected an fp64 vector of {len(STAT_KEYS)} statistics")
    if is_distributed():
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def stats_dict(vec: torch.Tensor) -> Dict[str, float]:
    return dict(zip(STAT_KEYS, vec.tolist()))


def max_over_ranks(value: float, device=None) -> float:
    """A timed region counts as its slowest rank."""
    if not is_distributed():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def global_seeds(seed: int, start: int, count: int, device=None) -> torch.Tensor:
    """Philox keys of the instances [start, start + count): ``seed + global id``."""
    return torch.arange(count, dtype=torch.int64, device=device) + int(seed) + int(start)


def fedavg_all_reduce(parameters: Dict[str, torch.Tensor], num_samples: float,
                      group=None, bucket_dtype=None) -> Dict[str, torch.Tensor]:
    """Sample-weighted average of per-rank client parameters - the reference's
    ``FedAvgAggregator.aggregate`` (federated/core.py:233-258: ``sum_i (n_i / N) p_i``) with the
    list of clients replaced by the ranks of a process group.  Every tensor is packed into one
    flat bucket (one ``cat``), scaled by the rank's sample count, moved by ONE all-reduce (NCCL over
    NVLink / NVSwitch on GPUs; the sample counts ride along as the bucket's last element) and handed
    back as views of the reduced bucket in the original shapes - a handful of kernels whatever the
    number of tensors.  The bucket has the parameters' own floating type when they all share one
    (fp32 for the reference's torch learners; its numpy arithmetic keeps float32 too), else fp64;
    ``bucket_dtype`` overrides.  A rank with ``num_samples == 0`` contributes nothing; if no rank
    has samples the result is empty, as upstream."""
    names = list(parameters)
    if not names:
        return {}
    tensors = [parameters[k] for k in names]
    dtypes = {t.dtype for t in tensors}
    if bucket_dtype is None:
        only = next(iter(dtypes)) if len(dtypes) == 1 else None
        bucket_dtype = only if only in (torch.float32, torch.float64) else torch.float64
    dev = tensors[0].device
    count = torch.full((1,), 1.0, dtype=bucket_dtype, device=dev)
    flat = torch.cat([t.reshape(-1).to(bucket_dtype) for t in tensors] + [count])
    flat *= float(num_samples)
    if is_distributed():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    total = float(flat[-1].item())
    if total == 0.0:
        return {}
    flat /= total
    out, o = {}, 0
    for k, t in zip(names, tensors):
        n = t.numel()
        v = flat[o:o + n].reshape(t.shape)
        out[k] = v if v.dtype == t.dtype else v.to(t.dtype)
        o += n
    return out
