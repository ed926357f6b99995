"""Partitioning of the hot path across the GPUs of one box (SURVEY.md section 8 e).

One process per GPU (`torch.distributed`).  Only three things shard, all without a
data-path collective except a final gather:

  * MSM            contiguous point range per rank, each rank a full Pippenger on its
                   slice, then a gather of ONE 96-byte Jacobian point per rank and a fold
                   (the same `results.iter().fold(identity, +)` upstream's best_multiexp
                   applies to its per-thread chunks).
  * column work    independent advice / permutation / lookup / quotient-piece NTTs and
                   commitments are dealt round-robin; nothing is exchanged.
  * h(X)           each rank evaluates a contiguous slice of the extended domain
                   (b200zk_quotient_env.range_*), the slices are all-gathered over NCCL.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous [begin, end) share of n items for `rank`; earlier ranks get the extras."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def assign_columns(n_columns: int, rank: int, world: int) -> list:
    """Round-robin ownership of independent columns."""
    return list(range(rank, n_columns, world))


def _dist():
    import torch.distributed as dist
    return dist


def all_gather_limbs(local: np.ndarray, device=None) -> np.ndarray:
    """All-gather a fixed-size uint64 array from every rank -> (world, *local.shape)."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[None, ...].copy()
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return torch.stack(outs).cpu().numpy().view(np.uint64)


def sharded_best_multiexp(coeffs: np.ndarray, bases: np.ndarray, local_msm=None, fold=None, device=None) -> np.ndarray:
    """`best_multiexp` over a point-range split.  Every rank passes the FULL (coeffs,
    bases) views (or at least its own slice filled in) and gets the full result.
    `local_msm` / `fold` default to the device path (best_multiexp / g1_sum)."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if local_msm is None or fold is None:
        from . import api
        local_msm = local_msm or api.best_multiexp
        fold = fold or api.g1_sum
    assert coeffs.shape[0] == bases.shape[0]
    b, e = shard_range(coeffs.shape[0], rank, world)
    partial = np.ascontiguousarray(local_msm(np.ascontiguousarray(coeffs[b:e]), np.ascontiguousarray(bases[b:e])))
    parts = all_gather_limbs(partial, device)
    return fold(np.ascontiguousarray(parts.reshape(world, 12)))


def gather_extended_chunks(local_chunk: np.ndarray, size: int, device=None) -> np.ndarray:
    """All-gather the per-rank slices of an extended-domain column (h(X)) into the full
    column.  Slices follow shard_range(size, rank, world) and may differ by one element,
    so they are padded to the longest slice for the collective."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        assert local_chunk.shape[0] == size
        return local_chunk
    longest = -(-size // world)
    padded = np.zeros((longest, 4), dtype=np.uint64)
    padded[: local_chunk.shape[0]] = local_chunk
    allc = all_gather_limbs(padded, device)
    out = np.zeros((size, 4), dtype=np.uint64)
    for r in range(world):
        b, e = shard_range(size, r, world)
        out[b:e] = allc[r, : e - b]
    return out
