"""The native witness generator for the reference's `pow_mod_fixed_exp` (reference
src/big_uint/chip.rs:454-490, `mul_mod` :355-413, `is_equal_muled` :513-608) against Python integers:
every value a mul_mod assigns — quotient, remainder, the carry-less limb products, the carries of the
equality check — and the final a^e mod n, for RSA-2048 in 32 limbs of 64 bits (reference src/lib.rs:266-268)
and for small shapes."""
import numpy as np
import pytest

import build as zkbuild


@pytest.fixture(scope="module")
def wit():
    zkbuild.build_witness()
    from b200zk import witness
    return witness


def _limbs(x: int, L: int) -> np.ndarray:
    return np.frombuffer(x.to_bytes(8 * L, "little"), dtype="<u8").copy()


def _int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype="<u8").tobytes(), "little")


def _expected_steps(a: int, n: int, e: int):
    acc, sq, steps = 1, a, []
    for bit in range(e.bit_length()):
        cur = sq
        steps.append((cur, cur))
        sq = cur * cur % n
        if (e >> bit) & 1:
            steps.append((acc, cur))
            acc = acc * cur % n
    return steps, acc


@pytest.mark.parametrize("L,e", [(32, 65537), (16, 65537), (4, 3), (2, 17), (1, 5)])
def test_pow_mod_fixed_exp_records_match_python_integers(wit, L, e):
    rnd = np.random.Generator(np.random.PCG64(1000 + L))
    count = 5
    ns, xs = [], []
    for i in range(count):
        n = int.from_bytes(rnd.bytes(8 * L), "little") | (1 << (64 * L - 1)) | 1
        x = int.from_bytes(rnd.bytes(8 * L), "little") % n
        if i == 1:
            x = n - 1
        if i == 2:
            x = 0
        ns.append(n)
        xs.append(x)
    recs, result = wit.pow_mod_fixed_exp(np.stack([_limbs(x, L) for x in xs]), np.stack([_limbs(n, L) for n in ns]), e, threads=3)
    M, base = 2 * L - 1, 1 << 64
    mx = L * (base - 1) ** 2 + (base - 1)                        # chip.rs:752-756
    for i in range(count):
        steps, want = _expected_steps(xs[i], ns[i], e)
        assert recs.shape[1] == len(steps) == e.bit_length() + bin(e).count("1")
        assert _int(result[i]) == want == pow(xs[i], e, ns[i])
        for s, (a, b) in enumerate(steps):
            r = wit.split_record(recs[i, s], L)
            assert _int(r.a) == a and _int(r.b) == b
            assert _int(r.q) == a * b // ns[i] and _int(r.r) == a * b % ns[i]
            al, bl, ql, nl, rl = ([int(v) for v in _limbs(z, L)] for z in (a, b, _int(r.q), ns[i], _int(r.r)))
            carry, extra = 0, 0
            for k in range(M):
                ab_k = sum(al[t] * bl[k - t] for t in range(max(0, k - L + 1), min(k, L - 1) + 1))
                qn_k = sum(ql[t] * nl[k - t] for t in range(max(0, k - L + 1), min(k, L - 1) + 1))
                assert _int(r.ab[k]) == ab_k and _int(r.qn[k]) == qn_k
                total = ab_k - (qn_k + (rl[k] if k < L else 0)) + carry + mx     # chip.rs:553-565
                assert total >= 0
                carry, c = divmod(total, base)
                extra += mx
                extra, mod_acc = divmod(extra, base)
                assert _int(r.carry[k]) == carry and int(r.c[k]) == c == mod_acc    # cs == mod_acc: a b = q n + r
            assert carry == extra


def test_base_not_below_modulus_is_rejected(wit):
    L = 4
    n = (1 << 255) | 12345
    with pytest.raises(Exception, match="base >= modulus"):
        wit.pow_mod_fixed_exp(_limbs(n, L)[None, :], _limbs(n, L)[None, :], 65537)


def test_words_to_fr_is_the_montgomery_form(wit):
    from oracle import bn254 as bn
    rnd = np.random.Generator(np.random.PCG64(5))
    vals = [0, 1, bn.R - 1, (1 << 133) - 7] + [int.from_bytes(rnd.bytes(24), "little") for _ in range(20)]
    words = np.stack([_limbs(v, 4) for v in vals])
    assert np.array_equal(wit.words_to_fr(words, threads=2), bn.fr_array_from_canonical(vals))
    w3 = np.stack([_limbs(v, 3) for v in vals if v < (1 << 192)])
    assert np.array_equal(wit.words_to_fr(w3), bn.fr_array_from_canonical([v for v in vals if v < (1 << 192)]))
