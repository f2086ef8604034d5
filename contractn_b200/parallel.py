"""Multi-GPU partitioning of the contraction path (one process per GPU).

The reference has no parallelism at all; these are the two places where the
path shards naturally:

* **batch sharding** -- independent network evaluations along a batch hub are
  split across ranks; each rank contracts its shard to a split-format pair
  ``(T_r, c_r)`` with its OWN scale, and :func:`merge_batch_shards` restores the
  reference's single global scale with two scalar all-reduces (the output
  stays sharded).
* **bond slicing** -- the largest contracted bonds are fixed to one value per
  slice, slices are distributed round-robin, every rank accumulates its slices
  with a log-sum-exp merge and :func:`merge_slice_sums` all-reduces the
  (mantissa, log-scale) pairs over NCCL/NVLink.

Both merges are exact by the gauge property of the split format (SURVEY.md
finding 0.4): ``T * exp(c)`` is what is summed / concatenated, the final pair is
re-canonicalised to ``mean|T| == 1``.  Works on CUDA tensors over NCCL and on
CPU tensors over gloo (the latter is how the logic is tested without GPUs).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def merge_batch_shards(T_local, c_local, group=None):
    """Global split-format pair for a batch-sharded evaluation.

    ``T_local`` is this rank's shard of the output (mean |T_local| == 1),
    ``c_local`` its log-scale: a 0-d/1-element float64 tensor (only element 0 is
    read).  Returns ``(T_shard, c)`` with ONE log-scale ``c`` shared by all ranks
    such that the concatenation of the shards has mean-abs 1, exactly what the
    reference returns for the unsharded network.
    """
    c_local = c_local.reshape(-1)[0].to(torch.float64)
    if _world(group) == 1:
        return T_local, c_local.reshape(1)
    c_star = c_local.clone()
    dist.all_reduce(c_star, op=dist.ReduceOp.MAX, group=group)
    w = torch.exp(c_local - c_star)
    n_local = float(T_local.numel())
    # T_local has mean-abs 1 by construction, so its abs-sum is numel * w after
    # bringing it to the common scale (no pass over the data needed)
    stats = torch.stack([w * n_local, torch.full_like(w, n_local)])
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    kappa = stats[0] / stats[1]
    T = T_local * (w / kappa).to(T_local.dtype)
    return T, (c_star + torch.log(kappa)).reshape(1)


def lse_accumulate(acc, T, c):
    """Running sum of split-format terms on one rank: ``acc`` is ``None`` or a
    pair ``(M, s)`` meaning ``M * exp(s)``; returns the pair including ``T*exp(c)``."""
    c = c.reshape(-1)[0].to(torch.float64)
    if acc is None:
        return T.clone(), c.clone()
    M, s = acc
    top = torch.maximum(s, c)
    M = M * torch.exp(s - top).to(M.dtype) + T * torch.exp(c - top).to(T.dtype)
    return M, top


def merge_slice_sums(M_local, s_local, group=None):
    """Sum over ranks of ``M_r * exp(s_r)`` returned in canonical split format
    ``(T, c)`` with mean |T| == 1 (all ranks get the full result)."""
    s_local = s_local.reshape(-1)[0].to(torch.float64)
    if _world(group) > 1:
        s_star = s_local.clone()
        dist.all_reduce(s_star, op=dist.ReduceOp.MAX, group=group)
        M = M_local * torch.exp(s_local - s_star).to(M_local.dtype)
        dist.all_reduce(M, op=dist.ReduceOp.SUM, group=group)
    else:
        s_star, M = s_local, M_local
    norm = M.abs().sum().to(torch.float64)
    numel = max(1, M.numel())
    if float(norm) > 1e-7:  # same threshold as stabilize (contractn/einsum.py:94-102)
        kappa = norm / numel
        return M / kappa.to(M.dtype), (s_star + torch.log(kappa)).reshape(1)
    return M, s_star.reshape(1)


def slice_assignments(n_slices, rank=None, world=None):
    """Round-robin slice indices of this rank."""
    if world is None:
        world = _world()
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    return list(range(rank, n_slices, world))
