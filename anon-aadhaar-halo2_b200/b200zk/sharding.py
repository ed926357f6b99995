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


def deal_by_load(total: int, per_item: list, fixed: list) -> list:
    """Contiguous [begin, end) ranges of `total` items, one per rank, such that fixed[r] + per_item[r] * count_r is
    as equal as integers allow (water filling): a rank with other work (cosets of h(X), the finish on rank 0) gets
    fewer columns to commit.  Ranks whose fixed cost exceeds the common finish time get none."""
    world = len(fixed)
    assert world == len(per_item) and world >= 1 and all(c > 0 for c in per_item)
    lo, hi = min(fixed), max(fixed) + max(per_item) * max(total, 1)
    for _ in range(64):                # the finish time t with sum_r max(0, (t - fixed_r) / per_item_r) = total
        t = 0.5 * (lo + hi)
        if sum(max(0.0, (t - fixed[r]) / per_item[r]) for r in range(world)) >= total:
            hi = t
        else:
            lo = t
    want = [max(0.0, (hi - fixed[r]) / per_item[r]) for r in range(world)]
    counts = [min(int(w), total) for w in want]
    short = total - sum(counts)
    order = sorted(range(world), key=lambda r: want[r] - counts[r], reverse=True)
    i = 0
    while short > 0:                   # hand the rounding remainder to the largest fractional parts
        counts[order[i % world]] += 1
        short -= 1
        i += 1
    while short < 0:
        r = max(range(world), key=lambda r: counts[r])
        counts[r] -= 1
        short += 1
    ranges, b = [], 0
    for r in range(world):
        ranges.append((b, b + counts[r]))
        b += counts[r]
    assert b == total
    return ranges


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


# ----------------------------------------------------------------- one NTT over several GPUs
# SURVEY.md section 8 e, "single large NTT": the four-step split n = n1 * n2 with ONE exchange.
# Input index j = j1 * n2 + j2, output index i = i1 + n1 * i2:
#   A[i1 + n1 i2] = sum_j2 w^(n1 i2 j2) * w^(i1 j2) * ( sum_j1 a[j1 n2 + j2] * w^(n2 i1 j1) )
# Rank r owns columns j2 in [r m, (r+1) m) (m = n2 / world) on the way in and rows
# i1 in [r n1/world, (r+1) n1/world) on the way out; neither is a contiguous slice of the
# natural order, which is what makes one all-to-all enough (the host-side scatter / gather of
# a natural-order vector is a strided copy, `column_block` / `natural_from_row_blocks`).

NTT_MAX_PASS_BITS = 9      # csrc/ntt.cuh NTT_MAX_B: one pass of the transform kernel covers at most 2^9 points


def _passes(bits: int) -> int:
    return max(1, -(-bits // NTT_MAX_PASS_BITS))


def four_step_split(k: int, world: int) -> int:
    """log2(n1) for a 2^k transform over `world` ranks (both factors >= world, n1 <= 2^9 so that the first step is one
    kernel pass): the split whose local transforms need the fewest kernel passes in total — k = 24: 2^8 x 2^16, 1 + 2
    passes, as many as the single-GPU transform (a balanced 2^12 x 2^12 needs 2 + 2) — and among those the balanced one
    up to k = 18, the one with 8-bit first passes (the single-GPU plan's width) above."""
    lw = world.bit_length() - 1
    assert world == 1 << lw, "world size must be a power of two"
    assert k >= 2 * lw and k >= 2, "transform too small for this many ranks"
    lo = max(lw, 1)
    hi = min(k - lo, NTT_MAX_PASS_BITS)
    assert hi >= lo, "transform too large for this many ranks"
    target = k // 2 if k <= 2 * NTT_MAX_PASS_BITS else 8
    return min(range(lo, hi + 1), key=lambda a: (_passes(a) + _passes(k - a), abs(a - target), a))


def column_block(a: np.ndarray, k: int, log_n1: int, world: int, rank: int) -> np.ndarray:
    """Rank's input share of a natural-order vector a (n, 4): its columns of the [n1][n2] matrix,
    X[j1][j2 - r m] = a[j1 n2 + j2]."""
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    m = n2 // world
    mat = a.reshape(n1, n2, 4)
    return np.ascontiguousarray(mat[:, rank * m:(rank + 1) * m])


def natural_from_row_blocks(blocks, k: int, log_n1: int) -> np.ndarray:
    """Inverse of the output distribution: blocks[s][il][i2] = A[(s n1/world + il) + n1 i2]."""
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    rows = np.concatenate(list(blocks), axis=0)            # [i1][i2]
    assert rows.shape[:2] == (n1, n2)
    return np.ascontiguousarray(rows.transpose(1, 0, 2)).reshape(n1 * n2, 4)


def sharded_best_fft(x_local, k: int, omega: int, ops, world: int, rank: int):
    """`best_fft` of 2^k elements over `world` ranks.  `x_local` is this rank's column block
    ([n1][m], see column_block); the result is its row block ([n1 / world][n2]).  `ops`
    supplies the local steps and the exchange (DeviceFourStep below on GPUs; the CPU tests
    inject the oracle):
        ops.first_pass_exchange(buf, k, log_n1, omega)  n1-point transforms down the columns, twiddle
                                                        omega^(i1 j2), exchange -> this rank's
                                                        [n1 / world][n2] row block
        ops.ntt_rows(buf, count, log_len, omega)        in-place transforms of contiguous rows
    """
    from .api import FR_MODULUS
    log_n1 = four_step_split(k, world)
    log_n2 = k - log_n1
    n1 = 1 << log_n1
    rows = ops.first_pass_exchange(x_local, k, log_n1, omega)
    ops.ntt_rows(rows, n1 // world, log_n2, pow(omega, n1, FR_MODULUS))
    return rows


class DeviceFourStep:
    """Device implementation of the two local steps and the exchange.  Buffers are torch
    int64 CUDA tensors (4 limbs per element).  Exchange modes:
      * "p2p"  : the first transform pass stores straight into the peers' row buffers through
                 symmetric-memory mappings (NVLink / NVSwitch): one kernel is local transform,
                 twiddle and all-to-all; no separate collective;
      * "nccl" : the same kernel packs per-destination chunks, `all_to_all_single`
                 moves them, a strided copy lays the rows out."""

    def __init__(self, k: int, world: int, rank: int, device, mode: str = "auto"):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import check, load
        self.C, self.torch, self.dist, self.check, self.lib = C, torch, dist, check, load()
        self.k, self.world, self.rank, self.device = k, world, rank, device
        self.log_n1 = four_step_split(k, world)
        n1, n2 = 1 << self.log_n1, 1 << (k - self.log_n1)
        self.rows_elems = (n1 // world) * n2
        # a stream of our own: handle 0 (torch's default stream) would mean "the library's
        # stream" to the C ABI, and the NCCL exchange must be ordered with the kernels
        self.stream = torch.cuda.Stream(device=device)
        self.mode = "nccl"
        self.rows = None
        if world > 1 and mode in ("auto", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm
                # two row buffers, used alternately: a peer may still be reading the rows of the previous
                # transform while this rank already stores into the other buffer, so the only barrier left is
                # the one after the stores ("my rows are complete")
                self.rows2 = [symm.empty(self.rows_elems * 4, dtype=torch.int64, device=device) for _ in range(2)]
                self.hdl2 = [symm.rendezvous(r, dist.group.WORLD.group_name) for r in self.rows2]
                self.peer_ptrs2 = [[int(p) for p in h.buffer_ptrs] for h in self.hdl2]
                self.turn = 0
                self.rows = self.rows2[0]
                self.mode = "p2p"
            except Exception as e:   # no peer mapping available: fall back to the NCCL exchange
                if mode == "p2p":
                    raise
                self.rows, self.p2p_error = None, repr(e)
        if self.rows is None:
            self.rows = torch.empty(self.rows_elems * 4, dtype=torch.int64, device=device)
        if self.mode == "nccl":
            self.send = torch.empty(self.rows_elems * 4, dtype=torch.int64, device=device)
            self.recv = torch.empty(self.rows_elems * 4, dtype=torch.int64, device=device) if world > 1 else None

    def _stream(self):
        return self.C.c_void_p(self.stream.cuda_stream)

    class _Ordered:
        """Run a block on self.stream, ordered after and before the caller's current stream."""

        def __init__(self, outer):
            self.o = outer

        def __enter__(self):
            t = self.o.torch
            self.cur = t.cuda.current_stream(self.o.device)
            self.o.stream.wait_stream(self.cur)
            self.ctx = t.cuda.stream(self.o.stream)
            self.ctx.__enter__()

        def __exit__(self, *exc):
            self.ctx.__exit__(*exc)
            self.cur.wait_stream(self.o.stream)
            return False

    def ntt_rows(self, buf, count: int, log_len: int, omega: int) -> None:
        from .api import _ptr, fr_limbs
        w = fr_limbs(omega)
        with self._Ordered(self):
            self.check(self.lib.b200zk_ntt_dev(self.C.c_void_p(buf.data_ptr()), 1 << log_len, count, log_len, _ptr(w),
                                               None, self._stream()))

    def first_pass_exchange(self, buf, k: int, log_n1: int, omega: int):
        with self._Ordered(self):
            return self._first_pass_exchange(buf, k, log_n1, omega)

    def _first_pass_exchange(self, buf, k: int, log_n1: int, omega: int):
        from .api import _ptr, fr_limbs
        C, world, rank = self.C, self.world, self.rank
        n1, n2 = 1 << log_n1, 1 << (k - log_n1)
        m, rows = n2 // world, n1 // world
        w = fr_limbs(omega)
        bases = (C.c_void_p * world)()
        if self.mode == "p2p":
            # buffer `turn` was last read two transforms ago, and every rank has passed the barrier of the transform
            # in between since then: it is free on every peer without another barrier
            turn = self.turn
            self.turn ^= 1
            for s in range(world):
                bases[s] = self.peer_ptrs2[turn][s]
            self.check(self.lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(buf.data_ptr()), k, log_n1, _ptr(w), world,
                                                                rank, bases, n2, rank * m, self._stream()))
            self.hdl2[turn].barrier(channel=0)   # every peer's stores into my rows have landed
            self.rows = self.rows2[turn]
            return self.rows
        dst = self.send if world > 1 else self.rows
        for s in range(world):
            bases[s] = dst.data_ptr() + s * rows * m * 32
        self.check(self.lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(buf.data_ptr()), k, log_n1, _ptr(w), world, rank,
                                                            bases, m, 0, self._stream()))
        if world == 1:
            return self.rows
        self.dist.all_to_all_single(self.recv, self.send)
        self.check(self.lib.b200zk_ntt4_gather_rows_dev(C.c_void_p(self.recv.data_ptr()), C.c_void_p(self.rows.data_ptr()),
                                                        k, log_n1, world, self._stream()))
        return self.rows
