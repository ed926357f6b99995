// Witness values of the reference's RSA modular exponentiation, computed natively and in parallel.
//
// once the prover's hot path runs in tens of milliseconds, the CPU
// time the reference's README reports (14.4 s of `MockProver::run`, src/lib.rs:947-959) is the
// Amdahl floor, and its arithmetic core is `BigUintConfig::pow_mod_fixed_exp`
// (reference src/big_uint/chip.rs:454-490): for e = 65537, seventeen `square_mod` and two `mul_mod`
// (chip.rs:355-413) over 2048-bit integers in 32 limbs of 64 bits, each of which assigns
//   * the quotient q and remainder r of a * b by n               (chip.rs:372-382),
//   * the carry-less limb products ab_k = sum_{i+j=k} a_i b_j and qn_k likewise (chip.rs:383-384,
//     `mul` :276-293 -> halo2-ecc mul_no_carry), and qn_k + r_k          (:387-407),
//   * the carries and low words of `is_equal_muled`'s running sum
//     sum_k = ab_k - (qn_k + r_k) + carry_k + max                (:513-608, max from :752-756).
// This file computes exactly those integers for a batch of independent (base, modulus) pairs, one
// thread per slice of the batch, with plain 64-bit limbs (unsigned __int128 products, Knuth's
// algorithm D for the division).  It is host code next to the C ABI, not part of the CUDA path; the
// values are what the circuit's `assign_integer` / `load_witness` calls would place in advice cells
// (canonical integers; b200zk_witness_words_to_fr gives the Montgomery `bn256::Fr` form).
//
// Layout of one modular multiplication record, in 64-bit words (L = num_limbs, M = 2L - 1):
//   a[L] b[L] q[L] r[L] ab[M][3] qn[M][3] carry[M][2] c[M]
// and one exponentiation = `steps` such records in the order the chip executes them, followed by
// the result x^e mod n [L].
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/b200zk_witness.h"

typedef unsigned __int128 u128;

namespace {

struct Wide { uint64_t w[3]; };   // a non-negative integer below 2^192

inline void wide_add_prod(Wide& acc, uint64_t x, uint64_t y) {
    const u128 p = (u128)x * y;
    u128 s = (u128)acc.w[0] + (uint64_t)p;
    acc.w[0] = (uint64_t)s;
    s = (u128)acc.w[1] + (uint64_t)(p >> 64) + (uint64_t)(s >> 64);
    acc.w[1] = (uint64_t)s;
    acc.w[2] += (uint64_t)(s >> 64);
}

// out[k] = sum_{i+j=k} x_i y_j, no carries between limbs (mul_no_carry)
void mul_no_carry(const uint64_t* x, const uint64_t* y, uint32_t L, Wide* out) {
    for (uint32_t k = 0; k < 2 * L - 1; ++k) out[k] = Wide{{0, 0, 0}};
    for (uint32_t i = 0; i < L; ++i)
        for (uint32_t j = 0; j < L; ++j) wide_add_prod(out[i + j], x[i], y[j]);
}

// prod[2L] = x * y with carries
void mul_full(const uint64_t* x, const uint64_t* y, uint32_t L, uint64_t* prod) {
    memset(prod, 0, 2 * L * sizeof(uint64_t));
    for (uint32_t i = 0; i < L; ++i) {
        uint64_t carry = 0;
        for (uint32_t j = 0; j < L; ++j) {
            const u128 t = (u128)x[i] * y[j] + prod[i + j] + carry;
            prod[i + j] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        prod[i + L] = carry;
    }
}

// Knuth, TAOCP vol. 2, 4.3.1 algorithm D in base 2^64: u[2L] / v[L] -> q[L + 1], r[L]; v != 0.
void div_rem(const uint64_t* u_in, const uint64_t* v_in, uint32_t L, uint64_t* q, uint64_t* r) {
    uint32_t n = L;
    while (n > 0 && v_in[n - 1] == 0) --n;
    const uint32_t m2 = 2 * L;
    std::vector<uint64_t> u(m2 + 1), v(n);
    for (uint32_t i = 0; i <= L; ++i) q[i] = 0;
    if (n == 1) {
        u128 rem = 0;
        std::vector<uint64_t> qq(m2);
        for (int i = (int)m2 - 1; i >= 0; --i) {
            const u128 cur = (rem << 64) | u_in[i];
            qq[i] = (uint64_t)(cur / v_in[0]);
            rem = cur % v_in[0];
        }
        for (uint32_t i = 0; i <= L && i < m2; ++i) q[i] = qq[i];
        memset(r, 0, L * sizeof(uint64_t));
        r[0] = (uint64_t)rem;
        return;
    }
    const int s = __builtin_clzll(v_in[n - 1]);
    for (uint32_t i = n; i-- > 0;) v[i] = s ? (v_in[i] << s) | (i ? v_in[i - 1] >> (64 - s) : 0) : v_in[i];
    u[m2] = s ? u_in[m2 - 1] >> (64 - s) : 0;
    for (uint32_t i = m2; i-- > 0;) u[i] = s ? (u_in[i] << s) | (i ? u_in[i - 1] >> (64 - s) : 0) : u_in[i];
    for (int j = (int)(m2 - n); j >= 0; --j) {
        const u128 num = ((u128)u[j + n] << 64) | u[j + n - 1];
        u128 qhat = num / v[n - 1], rhat = num % v[n - 1];
        while ((qhat >> 64) != 0 || qhat * v[n - 2] > ((rhat << 64) | u[j + n - 2])) {
            --qhat;
            rhat += v[n - 1];
            if ((rhat >> 64) != 0) break;
        }
        // multiply and subtract
        u128 borrow = 0, carry = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const u128 p = qhat * v[i] + carry;
            carry = p >> 64;
            const u128 t = (u128)u[i + j] - (uint64_t)p - borrow;
            u[i + j] = (uint64_t)t;
            borrow = (t >> 64) ? 1 : 0;
        }
        const u128 t = (u128)u[j + n] - carry - borrow;
        u[j + n] = (uint64_t)t;
        if (t >> 64) {   // qhat was one too large: add back
            --qhat;
            u128 c = 0;
            for (uint32_t i = 0; i < n; ++i) {
                const u128 a = (u128)u[i + j] + v[i] + c;
                u[i + j] = (uint64_t)a;
                c = a >> 64;
            }
            u[j + n] += (uint64_t)c;
        }
        if ((uint32_t)j <= L) q[j] = (uint64_t)qhat;
    }
    for (uint32_t i = 0; i < L; ++i)
        r[i] = i < n ? (s ? (u[i] >> s) | ((u128)u[i + 1] << (64 - s)) : u[i]) : 0;
}

// signed 256-bit helper for the running sum of is_equal_muled (values stay far below 2^200)
struct S256 { uint64_t w[4]; };
inline S256 s_from(const Wide& a) { return S256{{a.w[0], a.w[1], a.w[2], 0}}; }
inline S256 s_add(S256 a, const S256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.w[i] + b.w[i]; a.w[i] = (uint64_t)c; c >>= 64; }
    return a;
}
inline S256 s_sub(S256 a, const S256& b) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        const u128 t = (u128)a.w[i] - b.w[i] - br;
        a.w[i] = (uint64_t)t;
        br = (t >> 64) ? 1 : 0;
    }
    return a;
}

// one mul_mod record (chip.rs:355-413); returns false when is_equal_muled would fail (never for valid input)
bool mul_mod_record(const uint64_t* a, const uint64_t* b, const uint64_t* n, uint32_t L, uint64_t* rec) {
    const uint32_t M = 2 * L - 1;
    uint64_t* o_a = rec;
    uint64_t* o_b = o_a + L;
    uint64_t* o_q = o_b + L;
    uint64_t* o_r = o_q + L;
    uint64_t* o_ab = o_r + L;
    uint64_t* o_qn = o_ab + 3 * M;
    uint64_t* o_carry = o_qn + 3 * M;
    uint64_t* o_c = o_carry + 2 * M;
    memcpy(o_a, a, L * 8);
    memcpy(o_b, b, L * 8);
    std::vector<uint64_t> prod(2 * L), q(L + 1);
    mul_full(a, b, L, prod.data());
    div_rem(prod.data(), n, L, q.data(), o_r);
    memcpy(o_q, q.data(), L * 8);                 // a, b < n  =>  q < n: the top digit is zero
    std::vector<Wide> ab(M), qn(M);
    mul_no_carry(a, b, L, ab.data());
    mul_no_carry(o_q, n, L, qn.data());
    // muled_limb_max = min_n (2^64 - 1)^2 + (2^64 - 1), min_n = L            (chip.rs:752-756)
    S256 mx{{0, 0, 0, 0}};
    {
        const u128 m1 = ~(u128)0 >> 64;                    // 2^64 - 1
        const u128 sq = m1 * m1;                           // (2^64 - 1)^2 < 2^128
        u128 lo = 0, hi = 0;                               // L * sq as 192 bits
        for (uint32_t i = 0; i < L; ++i) { const u128 t = lo + sq; if (t < lo) ++hi; lo = t; }
        lo += m1;
        if (lo < m1) ++hi;
        mx.w[0] = (uint64_t)lo; mx.w[1] = (uint64_t)(lo >> 64); mx.w[2] = (uint64_t)hi;
    }
    S256 carry{{0, 0, 0, 0}}, extra{{0, 0, 0, 0}};
    bool ok = true;
    for (uint32_t k = 0; k < M; ++k) {
        memcpy(o_ab + 3 * k, ab[k].w, 24);
        memcpy(o_qn + 3 * k, qn[k].w, 24);
        S256 rhs = s_from(qn[k]);                          // qn_prod limb: qn_k + r_k for k < L
        if (k < L) rhs = s_add(rhs, S256{{o_r[k], 0, 0, 0}});
        // sum = ab_k - rhs + carry + max  (non-negative by construction of max)
        const S256 sum = s_add(s_add(s_sub(s_from(ab[k]), rhs), carry), mx);
        o_c[k] = sum.w[0];
        carry = S256{{sum.w[1], sum.w[2], sum.w[3], 0}};
        o_carry[2 * k] = carry.w[0];
        o_carry[2 * k + 1] = carry.w[1];
        extra = s_add(extra, mx);
        ok = ok && (sum.w[0] == extra.w[0]) && carry.w[2] == 0 && (sum.w[3] >> 63) == 0;
        extra = S256{{extra.w[1], extra.w[2], extra.w[3], 0}};
    }
    ok = ok && carry.w[0] == extra.w[0] && carry.w[1] == extra.w[1];   // the final carry equals accumulated_extra
    return ok;
}

uint32_t bit_length(uint64_t e) { return e ? 64 - (uint32_t)__builtin_clzll(e) : 0; }

}  // namespace

extern "C" {

size_t b200zk_witness_mul_mod_words(uint32_t num_limbs) {
    const size_t L = num_limbs, M = 2 * L - 1;
    return 4 * L + 3 * M + 3 * M + 2 * M + M;
}

uint32_t b200zk_witness_pow_steps(uint64_t e) {
    // chip.rs:465-488: one square_mod per bit of e, one mul_mod more per set bit
    return bit_length(e) + (uint32_t)__builtin_popcountll(e);
}

size_t b200zk_witness_pow_words(uint64_t e, uint32_t num_limbs) {
    return (size_t)b200zk_witness_pow_steps(e) * b200zk_witness_mul_mod_words(num_limbs) + num_limbs;
}

int b200zk_witness_pow_mod_fixed_exp(const uint64_t* base, const uint64_t* modulus, uint64_t e, uint32_t num_limbs,
                                     size_t count, int threads, uint64_t* out) {
    if (!base || !modulus || !out || num_limbs == 0 || num_limbs > 4096 || e == 0) return 1;
    const uint32_t L = num_limbs;
    const size_t rec = b200zk_witness_mul_mod_words(L), per = b200zk_witness_pow_words(e, L);
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(count, 1));
    std::vector<int> status(threads, 0);
    auto work = [&](int t) {
        const size_t lo = count * t / threads, hi = count * (t + 1) / threads;
        std::vector<uint64_t> acc(L), sq(L);
        for (size_t idx = lo; idx < hi; ++idx) {
            const uint64_t* a = base + idx * L;
            const uint64_t* n = modulus + idx * L;
            uint64_t* o = out + idx * per;
            bool zero_n = true, a_lt_n = false;
            for (uint32_t i = L; i-- > 0;) {
                if (n[i]) zero_n = false;
                if (a[i] != n[i]) { a_lt_n = a[i] < n[i]; break; }
            }
            if (zero_n || !a_lt_n) { status[t] = 2; continue; }   // the chip asserts x < n (src/chip.rs:88)
            std::fill(acc.begin(), acc.end(), 0);
            acc[0] = 1;                                            // assign_constant(1), extended with zero limbs
            memcpy(sq.data(), a, L * 8);
            const uint32_t nbits = bit_length(e);
            for (uint32_t bit = 0; bit < nbits; ++bit) {
                // `squared = square_mod(cur_sq)`, then `acc = mul_mod(acc, cur_sq)` when the bit is set
                if (!mul_mod_record(sq.data(), sq.data(), n, L, o)) status[t] = 3;
                const uint64_t* next_sq = o + 3 * L;               // r of this record
                o += rec;
                if ((e >> bit) & 1) {
                    if (!mul_mod_record(acc.data(), sq.data(), n, L, o)) status[t] = 3;
                    memcpy(acc.data(), o + 3 * L, L * 8);
                    o += rec;
                }
                memcpy(sq.data(), next_sq, L * 8);
            }
            memcpy(o, acc.data(), L * 8);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int s : status)
        if (s) return s;
    return 0;
}

// canonical little-endian integers of `width` words (< r) -> bn256::Fr Montgomery limbs (4 words each)
int b200zk_witness_words_to_fr(const uint64_t* words, uint32_t width, size_t count, int threads, uint64_t* out_fr) {
    if (!words || !out_fr || width == 0 || width > 4) return 1;
    static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static const uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
    const uint64_t INV = 0xc2e1f593efffffffull;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(count, 1));
    auto work = [&](int t) {
        for (size_t idx = count * t / threads; idx < count * (t + 1) / threads; ++idx) {
            uint64_t a[4] = {0, 0, 0, 0}, tt[6] = {0, 0, 0, 0, 0, 0};
            for (uint32_t i = 0; i < width; ++i) a[i] = words[idx * width + i];
            for (int i = 0; i < 4; ++i) {                          // CIOS: a * R^2 / R = a R mod p
                u128 c = 0;
                for (int j = 0; j < 4; ++j) { c += (u128)a[j] * R2[i] + tt[j]; tt[j] = (uint64_t)c; c >>= 64; }
                c += tt[4]; tt[4] = (uint64_t)c; tt[5] = (uint64_t)(c >> 64);
                const uint64_t m = tt[0] * INV;
                c = (u128)m * P[0] + tt[0]; c >>= 64;
                for (int j = 1; j < 4; ++j) { c += (u128)m * P[j] + tt[j]; tt[j - 1] = (uint64_t)c; c >>= 64; }
                c += tt[4]; tt[3] = (uint64_t)c; tt[4] = tt[5] + (uint64_t)(c >> 64);
            }
            bool ge = tt[4] != 0;
            if (!ge) {
                ge = true;
                for (int i = 3; i >= 0; --i)
                    if (tt[i] != P[i]) { ge = tt[i] > P[i]; break; }
            }
            if (ge) {
                u128 br = 0;
                for (int i = 0; i < 4; ++i) { const u128 d = (u128)tt[i] - P[i] - br; tt[i] = (uint64_t)d; br = (d >> 64) ? 1 : 0; }
            }
            memcpy(out_fr + 4 * idx, tt, 32);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"
