"""Host-side multi-GPU logic on the CPU: world_size-2 gloo processes exercise the
point-range MSM split, the one-point-per-rank gather + fold, the column deal and the
h(X) chunk gather.  The per-rank device work is replaced by the oracle here (this is a
test of the plumbing; the kernels themselves are covered by the -m gpu tests)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_range_covers_everything():
    from b200zk.sharding import assign_columns, shard_range

    for n in (0, 1, 7, 8, 1000, (1 << 17) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    cols = sorted(c for r in range(3) for c in assign_columns(11, r, 3))
    assert cols == list(range(11))


class _CpuFourStep:
    """Oracle stand-ins for the device steps of sharded_best_fft + a gloo all-to-all."""

    def __init__(self, dist, bn, co, world, rank):
        self.dist, self.bn, self.co, self.world, self.rank = dist, bn, co, world, rank

    def ntt_rows(self, buf, count, log_len, omega):
        w = self.bn.fr_array_from_canonical([omega])[0]
        for i in range(count):
            buf[i] = self.co.best_fft(np.ascontiguousarray(buf[i]), w, log_len, 1)

    def first_pass_exchange(self, buf, k, log_n1, omega):
        import torch
        bn, world, rank = self.bn, self.world, self.rank
        n1, n2 = 1 << log_n1, 1 << (k - log_n1)
        m, rows = n2 // world, n1 // world
        w1 = bn.fr_array_from_canonical([pow(omega, n2, bn.R)])[0]
        send = np.zeros((world, rows, m, 4), dtype=np.uint64)
        for jl in range(m):                                      # buf is the natural [n1][m] slab
            col = self.co.best_fft(np.ascontiguousarray(buf[:, jl]), w1, log_n1, 1)
            tw = [v * pow(omega, i1 * (rank * m + jl), bn.R) % bn.R for i1, v in enumerate(bn.fr_array_to_canonical(col))]
            send[:, :, jl] = bn.fr_array_from_canonical(tw).reshape(world, rows, 4)
        recv = torch.empty_like(torch.from_numpy(send.view(np.int64)))
        self.dist.all_to_all_single(recv, torch.from_numpy(send.view(np.int64)))
        recv = recv.numpy().view(np.uint64)                       # [r][il][jl]
        return np.ascontiguousarray(recv.transpose(1, 0, 2, 3)).reshape(rows, n2, 4)


def _four_step_check(rank, world, dist, sharding, bn, co):
    import torch
    k = 8
    a = co.gen_scalars(11, 1 << k)
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (28 - k), bn.R)
    log_n1 = sharding.four_step_split(k, world)
    x = sharding.column_block(a, k, log_n1, world, rank)
    rows = sharding.sharded_best_fft(x, k, omega, _CpuFourStep(dist, bn, co, world, rank), world, rank)
    t = torch.from_numpy(np.ascontiguousarray(rows).view(np.int64))
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    got = sharding.natural_from_row_blocks([o.numpy().view(np.uint64) for o in outs], k, log_n1)
    want = co.best_fft(a.copy(), bn.fr_array_from_canonical([omega])[0], k, 1)
    return bool(np.array_equal(got, want))


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
    import torch.distributed as dist

    from b200zk import sharding
    from oracle import bn254 as bn
    from oracle import c_oracle as co

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1001
        s, p = co.gen_scalars(7, n), co.gen_points(8, n)

        def fold(parts):   # oracle stand-in for b200zk_g1_sum
            acc = None
            for j in parts:
                acc = bn.g1_add(acc, bn.g1_jacobian_limbs_to_affine(j))
            return acc

        got = sharding.sharded_best_multiexp(s, p, local_msm=lambda a, b: co.best_multiexp(a, b, 2), fold=fold)
        exp = bn.g1_jacobian_limbs_to_affine(co.best_multiexp(s, p, 2))
        ok_msm = got == exp
        size = 515
        col = co.gen_scalars(9, size)
        b, e = sharding.shard_range(size, rank, world)
        full = sharding.gather_extended_chunks(col[b:e].copy(), size)
        ok_h = np.array_equal(full, col)
        ok_fft = _four_step_check(rank, world, dist, sharding, bn, co)
        q.put((rank, ok_msm, ok_h and ok_fft))
    finally:
        dist.destroy_process_group()


def test_four_step_split_minimises_kernel_passes():
    """Both factors at least the world size, the first one a single kernel pass (<= 2^9 points), the fewest passes
    in total: k = 24 is 2^8 x 2^16 (1 + 2 passes, as many as the single-GPU transform), small k stays k / 2."""
    from b200zk import sharding
    for world in (1, 2, 4, 8):
        lw = world.bit_length() - 1
        for k in range(max(2, 2 * lw), 29):
            a = sharding.four_step_split(k, world)
            assert a >= max(lw, 1) and k - a >= max(lw, 1)
            best = min(sharding._passes(b) + sharding._passes(k - b)
                       for b in range(max(lw, 1), min(k - max(lw, 1), sharding.NTT_MAX_PASS_BITS) + 1))
            assert sharding._passes(a) + sharding._passes(k - a) == best
            assert a <= sharding.NTT_MAX_PASS_BITS                 # the first step is one kernel pass
            if k <= 18:
                assert a == k // 2
    assert sharding.four_step_split(24, 8) == 8 and sharding.four_step_split(27, 8) == 9


def test_two_rank_gloo_msm_split_and_h_gather():
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results), results


def test_deal_by_load_equalises_finish_times():
    """The commitments of a sharded proof are dealt by load (b200zk/sharding.py deal_by_load): every item is dealt
    exactly once, in contiguous ranges, and fixed + per_item * count is level across ranks up to one item."""
    import random

    from b200zk.sharding import deal_by_load

    # the 8-GPU proof of DESIGN.md section 7: four coset owners (rank 0 also finishes h(X)), four plain ranks
    ranges = deal_by_load(242, [0.135] * 8, [3.4 + 1.05, 3.4, 3.4, 3.4, 0.0, 0.0, 0.0, 0.0])
    counts = [e - b for b, e in ranges]
    assert sum(counts) == 242 and ranges[0][0] == 0 and all(ranges[i][1] == ranges[i + 1][0] for i in range(7))
    assert counts[0] < min(counts[1:4]) and max(counts[1:4]) < min(counts[4:])
    assert max(counts[1:4]) - min(counts[1:4]) <= 1 and max(counts[4:]) - min(counts[4:]) <= 1
    fin = [f + 0.135 * c for f, c in zip([4.45, 3.4, 3.4, 3.4, 0, 0, 0, 0], counts)]
    assert max(fin) - min(fin) <= 0.135 + 1e-9
    # a rank whose other work exceeds everybody's finish time gets nothing; a single rank gets everything
    assert [e - b for b, e in deal_by_load(10, [1.0, 1.0], [100.0, 0.0])] == [0, 10]
    assert deal_by_load(7, [0.3], [5.0]) == [(0, 7)] and deal_by_load(0, [1.0, 2.0], [0.0, 0.0]) == [(0, 0), (0, 0)]
    rnd = random.Random(3)
    for _ in range(200):
        world = rnd.randint(1, 8)
        per = [rnd.uniform(0.05, 0.4) for _ in range(world)]
        fixed = [rnd.uniform(0.0, 6.0) for _ in range(world)]
        total = rnd.randint(0, 400)
        rg = deal_by_load(total, per, fixed)
        cnt = [e - b for b, e in rg]
        assert sum(cnt) == total and all(c >= 0 for c in cnt)
        busy = [r for r in range(world) if cnt[r] > 0]
        if busy:
            fin = [fixed[r] + per[r] * cnt[r] for r in busy]
            assert max(fin) - min(fin) <= 2 * max(per) + 1e-9
            assert all(fixed[r] >= min(fin) - 2 * max(per) for r in range(world) if cnt[r] == 0)
