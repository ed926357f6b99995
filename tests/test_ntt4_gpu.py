"""The four-step sharded NTT (b200zk_ntt4_* + sharding.sharded_best_fft) on one GPU: every
rank's steps are run one after the other in this process, with the fused first pass (local
transforms + twiddle + exchange in one kernel) storing into the other "ranks'" row buffers
exactly as it does into peer-mapped memory over NVLink.  The
result must equal the oracle's best_fft bit for bit.  (Two real ranks over gloo: see
tests/test_sharding_cpu.py; two real GPUs: scratch/gpu_ntt4.py under torchrun.)"""
import ctypes as C

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


class _OneGpuRanks:
    """`ops` for sharded_best_fft that plays rank `rank` of `world` on the local GPU."""

    def __init__(self, zk, k, world, rank, row_buffers):
        self.zk, self.lib, self.k, self.world, self.rank, self.rows = zk, zk.load(), k, world, rank, row_buffers

    def ntt_rows(self, col, count, log_len, omega):
        from b200zk.api import _ptr, fr_limbs
        w = fr_limbs(omega)
        self.zk.check(self.lib.b200zk_ntt_dev(C.c_void_p(col.ptr), 1 << log_len, count, log_len, _ptr(w), None, None))

    def first_pass_exchange(self, col, k, log_n1, omega):
        from b200zk.api import _ptr, fr_limbs
        n2 = 1 << (k - log_n1)
        m = n2 // self.world
        bases = (C.c_void_p * self.world)(*[r.ptr for r in self.rows])
        w = fr_limbs(omega)
        self.zk.check(self.lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(col.ptr), k, log_n1, _ptr(w), self.world,
                                                                  self.rank, bases, n2, self.rank * m, None))
        return None


@pytest.mark.parametrize("k,world", [(4, 1), (6, 2), (9, 2), (10, 4), (13, 8), (16, 4), (19, 4), (20, 8), (21, 2)])
def test_four_step_matches_best_fft(zk, k, world):
    from b200zk import sharding
    from b200zk.api import FR_MODULUS, FR_ROOT_OF_UNITY
    n = 1 << k
    a = co.gen_scalars(0x4E54 + k, n)
    omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
    want = co.best_fft(a.copy(), bn.fr_array_from_canonical([omega])[0], k, 2)
    log_n1 = sharding.four_step_split(k, world)
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    rows = [zk.DeviceColumn((n1 // world) * n2) for _ in range(world)]
    cols = [zk.DeviceColumn.from_host(sharding.column_block(a, k, log_n1, world, r).reshape(-1, 4)) for r in range(world)]
    lib = zk.load()
    # step 1 of every rank (the exchange is complete once all of them have stored)
    for r in range(world):
        _OneGpuRanks(zk, k, world, r, rows).first_pass_exchange(cols[r], k, log_n1, omega)
    # step 2 of every rank
    blocks = []
    for r in range(world):
        _OneGpuRanks(zk, k, world, r, rows).ntt_rows(rows[r], n1 // world, k - log_n1, pow(omega, n1, FR_MODULUS))
        blocks.append(rows[r].to_host().reshape(n1 // world, n2, 4))
    got = sharding.natural_from_row_blocks(blocks, k, log_n1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("k,world", [(6, 2), (12, 4), (16, 8)])
def test_fused_first_pass_equals_unfused_steps(zk, k, world):
    """b200zk_ntt4_first_pass_scatter_dev on the natural [n1][m] slab stores exactly what b200zk_ntt_dev +
    b200zk_ntt4_twiddle_scatter_dev store from the column-major [m][n1] slab (values leave the unfused
    kernel canonical, the fused one Montgomery-reduced only: compared after the row transforms)."""
    from b200zk import sharding
    from b200zk.api import FR_MODULUS, FR_ROOT_OF_UNITY, _ptr, fr_limbs
    n = 1 << k
    a = co.gen_scalars(0x4E77 + k, n)
    omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
    log_n1 = sharding.four_step_split(k, world)
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    m = n2 // world
    lib = zk.load()
    w, w1, w2 = fr_limbs(omega), fr_limbs(pow(omega, n2, FR_MODULUS)), fr_limbs(pow(omega, n1, FR_MODULUS))
    out = []
    for fused in (True, False):
        rows = [zk.DeviceColumn((n1 // world) * n2) for _ in range(world)]
        bases = (C.c_void_p * world)(*[r.ptr for r in rows])
        for r in range(world):
            slab = sharding.column_block(a, k, log_n1, world, r)                      # [n1][m]
            if fused:
                col = zk.DeviceColumn.from_host(slab.reshape(-1, 4))
                zk.check(lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(col.ptr), k, log_n1, _ptr(w), world, r, bases, n2, r * m, None))
            else:
                col = zk.DeviceColumn.from_host(np.ascontiguousarray(slab.transpose(1, 0, 2)).reshape(-1, 4))
                zk.check(lib.b200zk_ntt_dev(C.c_void_p(col.ptr), n1, m, log_n1, _ptr(w1), None, None))
                zk.check(lib.b200zk_ntt4_twiddle_scatter_dev(C.c_void_p(col.ptr), k, log_n1, _ptr(w), world, r, bases, n2, r * m, None))
        for r in range(world):
            zk.check(lib.b200zk_ntt_dev(C.c_void_p(rows[r].ptr), n2, n1 // world, k - log_n1, _ptr(w2), None, None))
        out.append(sharding.natural_from_row_blocks([rw.to_host().reshape(n1 // world, n2, 4) for rw in rows], k, log_n1))
    assert np.array_equal(out[0], out[1])
    assert np.array_equal(out[0], co.best_fft(a.copy(), bn.fr_array_from_canonical([omega])[0], k, 2))


def test_packed_exchange_layout(zk):
    """The NCCL form: pack per destination, then b200zk_ntt4_gather_rows_dev; world = 1 makes the
    exchange the identity, so pack + gather must reproduce the direct row block.  (The unfused
    kernels on a column-major slab.)"""
    from b200zk import sharding
    from b200zk.api import FR_MODULUS, FR_ROOT_OF_UNITY, _ptr, fr_limbs
    k, world = 11, 1
    n = 1 << k
    a = co.gen_scalars(0x4E99, n)
    omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
    log_n1 = sharding.four_step_split(k, world)
    n1, n2 = 1 << log_n1, 1 << (k - log_n1)
    lib = zk.load()
    col = zk.DeviceColumn.from_host(np.ascontiguousarray(sharding.column_block(a, k, log_n1, world, 0).transpose(1, 0, 2)).reshape(-1, 4))
    send, rows = zk.DeviceColumn(n), zk.DeviceColumn(n)
    w = fr_limbs(omega)
    w1, w2 = fr_limbs(pow(omega, n2, FR_MODULUS)), fr_limbs(pow(omega, n1, FR_MODULUS))   # kept alive across the calls
    zk.check(lib.b200zk_ntt_dev(C.c_void_p(col.ptr), n1, n2, log_n1, _ptr(w1), None, None))
    bases = (C.c_void_p * 1)(send.ptr)
    zk.check(lib.b200zk_ntt4_twiddle_scatter_dev(C.c_void_p(col.ptr), k, log_n1, _ptr(w), 1, 0, bases, n2, 0, None))
    zk.check(lib.b200zk_ntt4_gather_rows_dev(C.c_void_p(send.ptr), C.c_void_p(rows.ptr), k, log_n1, 1, None))
    zk.check(lib.b200zk_ntt_dev(C.c_void_p(rows.ptr), n2, n1, k - log_n1, _ptr(w2), None, None))
    got = sharding.natural_from_row_blocks([rows.to_host().reshape(n1, n2, 4)], k, log_n1)
    want = co.best_fft(a.copy(), bn.fr_array_from_canonical([omega])[0], k, 2)
    assert np.array_equal(got, want)


def test_bad_split_is_rejected(zk):
    from b200zk.api import _ptr, fr_limbs
    lib = zk.load()
    col = zk.DeviceColumn(16)
    bases = (C.c_void_p * 1)(col.ptr)
    w = fr_limbs(1)
    rc = lib.b200zk_ntt4_twiddle_scatter_dev(C.c_void_p(col.ptr), 4, 2, _ptr(w), 3, 0, bases, 4, 0, None)
    assert rc != 0 and b"world" in lib.b200zk_last_error()
    rc = lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(col.ptr), 4, 2, _ptr(w), 3, 0, bases, 4, 0, None)
    assert rc != 0 and b"world" in lib.b200zk_last_error()
    rc = lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(col.ptr), 24, 12, _ptr(w), 1, 0, bases, 4096, 0, None)
    assert rc != 0 and b"one pass" in lib.b200zk_last_error()
