"""World / pair sharding across the GPUs of one box and the (only) collectives of the path.

The reference is single-process, single-GPU (train/env_utils.py:18,26 hard-codes gpu 0; SURVEY.md
section 2.4).  Worlds are independent, so the multi-GPU form is: one process per GPU
(torchrun), contiguous world ranges (or cross-play pair slices) per rank, replicated policy
weights, NO collective inside the step loop.  NCCL (gloo in the CPU tests) is used only to

* sum per-slice (episode return, episode count) statistics -> mean returns
  (replaces ``scores.extend(running_score[dones].tolist())``, train/MAPPO/main_player.py:256-261,
  and the per-slice ``xp_scores`` of train/XD/xd_player.py:143-149), and
* assemble the ``[n, n]`` cross-play return matrix (BASELINE config 5).

Payloads are a few KB once per rollout.  Every function here works on CPU tensors with the
``gloo`` backend and on CUDA tensors with ``nccl``; without an initialised process group they
are the identity (single GPU).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def world_shard(n_worlds_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first world, number of worlds) of `rank`: contiguous, sizes differ by at most one"""
    base, rem = divmod(n_worlds_total, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_pairs(n_policies: int) -> List[Tuple[int, int]]:
    """ordered (seat-0 policy, seat-1 policy) pairs, row-major: pair k = (k // n, k % n)"""
    return [(i, j) for i in range(n_policies) for j in range(n_policies)]


def pair_shard(pairs: Sequence[Tuple[int, int]], rank: int, world_size: int) -> List[Tuple[int, int]]:
    """contiguous block of the pair list evaluated by `rank` (32 of the 256 pairs per GPU in config 5)"""
    first, count = world_shard(len(pairs), rank, world_size)
    return list(pairs[first:first + count])


def reduce_episode_stats(return_sum: torch.Tensor, episodes: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """global (sum of episode returns, number of episodes) from per-world device accumulators:
    one all-reduce of 2 int64"""
    v = torch.stack([return_sum.sum().to(torch.int64), episodes.sum().to(torch.int64)])
    if _world(group)[1] > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v[0], v[1]


def gather_pair_matrix(pairs_local: Sequence[Tuple[int, int]], return_sum: torch.Tensor, episodes: torch.Tensor,
                       n_policies: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Assemble the cross-play matrix from the per-pair statistics of every rank.

    Each rank contributes ``(i, j, return_sum, episodes)`` rows for the pairs it evaluated;
    ranks may hold different numbers of pairs (rows are padded to the longest shard for the
    all-gather).  Returns ``(mean_return float64 [n, n], episodes int64 [n, n])``; pairs nobody
    evaluated are NaN / 0.  A pair evaluated by several ranks (world-sharded pairs) is pooled."""
    rank, world = _world(group)
    dev = return_sum.device
    k = len(pairs_local)
    if return_sum.numel() != k or episodes.numel() != k:
        raise ValueError("one (return_sum, episodes) entry per local pair expected")
    rows = torch.empty((k, 4), dtype=torch.int64, device=dev)
    if k:
        rows[:, :2] = torch.as_tensor(list(pairs_local), dtype=torch.int64, device=dev).reshape(k, 2)
        rows[:, 2] = return_sum.to(torch.int64)
        rows[:, 3] = episodes.to(torch.int64)
    if world > 1:
        counts = [torch.zeros((1,), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([k], dtype=torch.int64, device=dev), group=group)
        kmax = int(max(int(c.item()) for c in counts))
        padded = torch.full((kmax, 4), -1, dtype=torch.int64, device=dev)
        padded[:k] = rows
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        rows = torch.cat([p[: int(c.item())] for p, c in zip(parts, counts)])
    sums = torch.zeros((n_policies, n_policies), dtype=torch.int64, device=dev)
    eps = torch.zeros_like(sums)
    if rows.numel():
        if int(rows[:, :2].min()) < 0 or int(rows[:, :2].max()) >= n_policies:
            raise ValueError("pair index out of range")
        flat = rows[:, 0] * n_policies + rows[:, 1]
        sums.view(-1).index_add_(0, flat, rows[:, 2])
        eps.view(-1).index_add_(0, flat, rows[:, 3])
    mean = sums.to(torch.float64) / eps.to(torch.float64)  # 0/0 -> NaN for pairs nobody played
    return mean, eps
