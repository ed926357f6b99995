// BN254 Fr / Fq Montgomery arithmetic on 8 x 32-bit limbs for sm_100a.
//
// Replaces (on the device) halo2curves 0.3.1 `bn256::Fr` / `bn256::Fq`
// ([DEP] halo2curves/src/bn256/{fr,fq}.rs, pinned at reference Cargo.lock:484-486):
// same Montgomery radix R = 2^256, so the 4 x u64 little-endian limbs the Rust side
// holds are reinterpreted as 8 x u32 with no conversion.
//
// Multiplication is a CIOS Montgomery product split into an "even" and an "odd"
// accumulator so that every partial product is one 64-bit multiply-accumulate
// (mad.lo.cc + madc.hi.cc, which ptxas fuses into IMAD.WIDE.U32 with carry) and no
// carry ever has to ripple between misaligned columns.  After each row the implicit
// division by 2^32 swaps the roles of the two accumulators; with the loops fully
// unrolled the swap and the two-limb shift are pure register renaming.
//
// Every function is __host__ __device__: the host bodies are a plain-C emulation of
// the same limb schedule so the algorithm is unit-tested on the CPU (tests/
// test_field_host.py) before any GPU time is spent.  In the product the host bodies
// only ever prepare per-call scalar kernel parameters (e.g. zeta^2 or n^-1 * zeta);
// array data is never processed on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zk {

// ------------------------------------------------------------------ field parameters
struct FrParams {
    // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    // (reference solidity_verifier_contract/contract.sol:211)
    static constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xefffffffu;  // -r^{-1} mod 2^32
    // R mod r
    static constexpr uint32_t ONE0 = 0x4ffffffbu, ONE1 = 0xac96341cu, ONE2 = 0x9f60cd29u, ONE3 = 0x36fc7695u,
                              ONE4 = 0x7879462eu, ONE5 = 0x666ea36fu, ONE6 = 0x9a07df2fu, ONE7 = 0x0e0a77c1u;
    // R^2 mod r
    static constexpr uint32_t RR0 = 0xae216da7u, RR1 = 0x1bb8e645u, RR2 = 0xe35c59e3u, RR3 = 0x53fe3ab1u,
                              RR4 = 0x53bb8085u, RR5 = 0x8c49833du, RR6 = 0x7f4e44a5u, RR7 = 0x0216d0b1u;
};

struct FqParams {
    // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    // (reference solidity_verifier_contract/contract.sol:210)
    static constexpr uint32_t P0 = 0xd87cfd47u, P1 = 0x3c208c16u, P2 = 0x6871ca8du, P3 = 0x97816a91u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t INV = 0xe4866389u;
    static constexpr uint32_t ONE0 = 0xc58f0d9du, ONE1 = 0xd35d438du, ONE2 = 0xf5c70b3du, ONE3 = 0x0a78eb28u,
                              ONE4 = 0x7879462cu, ONE5 = 0x666ea36fu, ONE6 = 0x9a07df2fu, ONE7 = 0x0e0a77c1u;
    static constexpr uint32_t RR0 = 0x538afa89u, RR1 = 0xf32cfc5bu, RR2 = 0xd44501fbu, RR3 = 0xb5e71911u,
                              RR4 = 0x0a417ff6u, RR5 = 0x47ab1effu, RR6 = 0xcab8351fu, RR7 = 0x06d89f71u;
};

// -------------------------------------------------------------- limb-chain primitives
// acc(8 limbs) += sum_t x[t] * y * 2^(64 t), t = 0..3; returns the carry out of limb 7.
// `cin` (0/1) is added at limb 0.
ZK_HD uint32_t mad_row4(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    return c;
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t carry = 0;
    for (int t = 0; t < 4; ++t) {
        unsigned __int128 s = (unsigned __int128)x[t] * y + (((uint64_t)acc[2 * t + 1] << 32) | acc[2 * t]) + carry;
        acc[2 * t] = (uint32_t)s;
        acc[2 * t + 1] = (uint32_t)(s >> 32);
        carry = (uint64_t)(s >> 64);
    }
    return (uint32_t)carry;
#endif
}

// lo += stray (one limb); the carry of that add enters the chain acc += x*y as above.
// No carry can leave limb 7 here (see the bound argument in mont_mul).
ZK_HD void mad_row4_stray(uint32_t& lo, uint32_t stray, uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2,
                          uint32_t x3, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(lo)
        : "r"(stray), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    uint64_t s0 = (uint64_t)lo + stray;
    lo = (uint32_t)s0;
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t carry = s0 >> 32;
    for (int t = 0; t < 4; ++t) {
        unsigned __int128 s = (unsigned __int128)x[t] * y + (((uint64_t)acc[2 * t + 1] << 32) | acc[2 * t]) + carry;
        acc[2 * t] = (uint32_t)s;
        acc[2 * t + 1] = (uint32_t)(s >> 32);
        carry = (uint64_t)(s >> 64);
    }
#endif
}

ZK_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }

// acc(2N limbs) += sum_{t<N} x[2t] * y * 2^(64 t); the carry out of limb 2N-1 is added to acc[2N], which the
// callers arrange to hold nothing but earlier such carries (see sqr()).  x is read with stride 2.
template <int N> ZK_HD void mad_chain(uint32_t* acc, const uint32_t* x, uint32_t y) {
#ifdef __CUDA_ARCH__
    if constexpr (N == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(x[0]), "r"(y));
    } else if constexpr (N == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(x[0]), "r"(x[2]), "r"(y));
    } else if constexpr (N == 3) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
            : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(y));
    } else {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
              "+r"(acc[7]), "+r"(acc[8])
            : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(y));
    }
#else
    uint64_t carry = 0;
    for (int t = 0; t < N; ++t) {
        unsigned __int128 s = (unsigned __int128)x[2 * t] * y + (((uint64_t)acc[2 * t + 1] << 32) | acc[2 * t]) + carry;
        acc[2 * t] = (uint32_t)s;
        acc[2 * t + 1] = (uint32_t)(s >> 32);
        carry = (uint64_t)(s >> 64);
    }
    acc[2 * N] += (uint32_t)carry;
#endif
}

// r(16 limbs) = a + b (16 limbs each); no carry leaves limb 15 for the operands sqr() forms
ZK_HD void add16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %16, %32;\n\t"
        "addc.cc.u32 %1, %17, %33;\n\t"
        "addc.cc.u32 %2, %18, %34;\n\t"
        "addc.cc.u32 %3, %19, %35;\n\t"
        "addc.cc.u32 %4, %20, %36;\n\t"
        "addc.cc.u32 %5, %21, %37;\n\t"
        "addc.cc.u32 %6, %22, %38;\n\t"
        "addc.cc.u32 %7, %23, %39;\n\t"
        "addc.cc.u32 %8, %24, %40;\n\t"
        "addc.cc.u32 %9, %25, %41;\n\t"
        "addc.cc.u32 %10, %26, %42;\n\t"
        "addc.cc.u32 %11, %27, %43;\n\t"
        "addc.cc.u32 %12, %28, %44;\n\t"
        "addc.cc.u32 %13, %29, %45;\n\t"
        "addc.cc.u32 %14, %30, %46;\n\t"
        "addc.u32 %15, %31, %47;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]),
          "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
          "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]),
          "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 16; ++i) {
        uint64_t s = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// lo += stray, returning the carry (0 / 1) of that addition
ZK_HD uint32_t add_carry_out(uint32_t& lo, uint32_t stray) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, 0, 0;"
        : "+r"(lo), "=r"(c) : "r"(stray));
    return c;
#else
    uint64_t s = (uint64_t)lo + stray;
    lo = (uint32_t)s;
    return (uint32_t)(s >> 32);
#endif
}

// acc(8 limbs) += sum_t x[t] * y * 2^(64 t) + cin (cin = 0 / 1 enters at limb 0); no carry leaves limb 7
ZK_HD void mad_row4_cin(uint32_t* acc, uint32_t cin, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "add.cc.u32 t, %8, 0xffffffff;\n\t"          // carry flag = cin
        "madc.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.u32 %7, %12, %13, %7;\n\t"
        "}"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(cin), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t carry = cin;
    for (int t = 0; t < 4; ++t) {
        unsigned __int128 s = (unsigned __int128)x[t] * y + (((uint64_t)acc[2 * t + 1] << 32) | acc[2 * t]) + carry;
        acc[2 * t] = (uint32_t)s;
        acc[2 * t + 1] = (uint32_t)(s >> 32);
        carry = (uint64_t)(s >> 64);
    }
#endif
}


// r = a + b (8 limbs), returns carry
ZK_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t s = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)s;
        c = s >> 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b (8 limbs), returns borrow (1 if a < b)
ZK_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;  // subc.u32 0-0-borrow = 0xffffffff when borrow
#else
    int64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        int64_t s = (int64_t)a[i] - b[i] - c;
        r[i] = (uint32_t)s;
        c = (s < 0) ? 1 : 0;
    }
    return (uint32_t)c;
#endif
}

// ------------------------------------------------------------------------ the field
template <class P>
struct alignas(16) Fp {
    using Params = P;
    uint32_t l[8];

    static ZK_HD void modulus(uint32_t* m) {
        m[0] = P::P0; m[1] = P::P1; m[2] = P::P2; m[3] = P::P3;
        m[4] = P::P4; m[5] = P::P5; m[6] = P::P6; m[7] = P::P7;
    }
    static ZK_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = 0;
        return r;
    }
    static ZK_HD Fp one() {  // Montgomery form of 1
        Fp r;
        r.l[0] = P::ONE0; r.l[1] = P::ONE1; r.l[2] = P::ONE2; r.l[3] = P::ONE3;
        r.l[4] = P::ONE4; r.l[5] = P::ONE5; r.l[6] = P::ONE6; r.l[7] = P::ONE7;
        return r;
    }
    static ZK_HD Fp r2() {
        Fp r;
        r.l[0] = P::RR0; r.l[1] = P::RR1; r.l[2] = P::RR2; r.l[3] = P::RR3;
        r.l[4] = P::RR4; r.l[5] = P::RR5; r.l[6] = P::RR6; r.l[7] = P::RR7;
        return r;
    }

    static ZK_HD void modulus2(uint32_t* m) {  // 2p
        m[0] = P::P0 << 1;
        m[1] = (P::P1 << 1) | (P::P0 >> 31); m[2] = (P::P2 << 1) | (P::P1 >> 31);
        m[3] = (P::P3 << 1) | (P::P2 >> 31); m[4] = (P::P4 << 1) | (P::P3 >> 31);
        m[5] = (P::P5 << 1) | (P::P4 >> 31); m[6] = (P::P6 << 1) | (P::P5 >> 31);
        m[7] = (P::P7 << 1) | (P::P6 >> 31);
    }

    // LAZY REDUCTION: every value of this type lives in [0, 2p) (p < 2^254, so 2p < 2^255
    // and 4p < 2^256 = R).  Arithmetic keeps that invariant without ever reducing to
    // [0, p); `canon()` produces the unique representative and is applied wherever a
    // value leaves the library (stores to user-visible memory, digit extraction).
    ZK_HD bool is_zero() const {   // value == 0 mod p, i.e. the limbs are 0 or p
        const uint32_t z = l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7];
        const uint32_t e = (l[0] ^ P::P0) | (l[1] ^ P::P1) | (l[2] ^ P::P2) | (l[3] ^ P::P3) |
                           (l[4] ^ P::P4) | (l[5] ^ P::P5) | (l[6] ^ P::P6) | (l[7] ^ P::P7);
        return z == 0 || e == 0;
    }
    ZK_HD bool operator==(const Fp& o) const { return (*this - o).is_zero(); }
    ZK_HD bool operator!=(const Fp& o) const { return !(*this == o); }
    // bit-for-bit comparison of representatives (host-side cache keys)
    ZK_HD bool same_limbs(const Fp& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) d |= l[i] ^ o.l[i];
        return d == 0;
    }

    // value in [0, 2p) -> [0, p)
    ZK_HD void reduce_once() {
        uint32_t m[8], t[8];
        modulus(m);
        uint32_t borrow = sub8(t, l, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = borrow ? l[i] : t[i];
    }
    ZK_HD Fp canon() const {
        Fp r = *this;
        r.reduce_once();
        return r;
    }

    friend ZK_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t m[8], t[8];
        modulus2(m);
        add8(r.l, a.l, b.l);  // a + b < 4p < 2^256: no carry
        uint32_t borrow = sub8(t, r.l, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = borrow ? r.l[i] : t[i];
        return r;
    }
    friend ZK_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t m[8], t[8];
        modulus2(m);
        uint32_t borrow = sub8(r.l, a.l, b.l);   // in (-2p, 2p)
        add8(t, r.l, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = borrow ? t[i] : r.l[i];
        return r;
    }
    ZK_HD Fp neg() const { return zero() - *this; }
    ZK_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*2^-256 mod p; inputs and output in [0, 2p): with 4p < R the
    // result (a*b + M*p)/R < 4p^2/R + p < 2p needs no final subtraction.
    //
    // T = E + 2^32 * O (+ one stray limb carried between rows).  Row i adds a*b_i and
    // m*p with m chosen so the low limb cancels; V = T + a*b_i + m*p < 2^288, hence
    // O (weight 2^32) never overflows 8 limbs and E needs one carry limb E[8].
    // Dividing by 2^32 turns O into the next row's even accumulator unchanged, E[2..8]
    // into the next odd accumulator, and leaves E[1] as a stray limb of weight 1 that
    // is added to the new E[0], its carry entering the new O chain (same weight 2^32).
    friend ZK_HD Fp operator*(const Fp& a, const Fp& b) {
        uint32_t m[8];
        modulus(m);
        uint32_t stray = 0;
        uint32_t E[9], O[9];  // E[8] is the carry limb of the even accumulator
#pragma unroll
        for (int i = 0; i < 9; ++i) { E[i] = 0; O[i] = 0; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t bi = b.l[i];
            mad_row4_stray(E[0], stray, O, a.l[1], a.l[3], a.l[5], a.l[7], bi);
            E[8] += mad_row4(E, a.l[0], a.l[2], a.l[4], a.l[6], bi);
            const uint32_t mi = mul_lo(E[0], P::INV);
            E[8] += mad_row4(E, m[0], m[2], m[4], m[6], mi);
            mad_row4(O, m[1], m[3], m[5], m[7], mi);
            // divide by 2^32: rename
            stray = E[1];
            uint32_t nO[9];
#pragma unroll
            for (int j = 0; j < 7; ++j) nO[j] = E[j + 2];
            nO[7] = 0; nO[8] = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) E[j] = O[j];
            E[8] = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) O[j] = nO[j];
        }
        // merge: T = E + stray + 2^32 * O   (T < 2p < 2^255)
        Fp r;
        uint32_t s[8];
        s[0] = stray;
#pragma unroll
        for (int j = 1; j < 8; ++j) s[j] = O[j - 1];
        add8(r.l, E, s);
        return r;
    }
    // Montgomery square a*a*2^-256 mod p, the same representative operator* returns (the reduction multiplier
    // M = -T p^-1 mod 2^256 depends on T = a^2 only), with 36 + 64 + 8 multiplier-pipe operations instead of
    // 64 + 64 + 8:
    //   1. the 28 cross products a_i a_j (i < j) once, in two 16-limb accumulators like operator*'s: a product
    //      at limb position i + j goes to E (even positions) or to O (odd positions, O[k] = limb k + 1), so every
    //      chain is aligned multiply-adds; rows are ordered so that a chain's final carry lands in a limb that
    //      holds only earlier carries;
    //   2. T = 2 (E + 2^32 O) + sum_i a_i^2 2^(64 i)   (shifts and adds: the other pipe);
    //   3. Q = (T mod R + M p) / R by eight reduction rows (operator*'s rows without the a b_i terms), and the
    //      result Q + floor(T / R) < 2p.
    ZK_HD Fp sqr() const {
#ifdef B200ZK_SQR_BY_MUL           // A/B builds only (build.py --variant)
        return (*this) * (*this);
#endif
        uint32_t E[17], O[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) { E[i] = 0; O[i] = 0; }
        const uint32_t* a = l;
        // odd positions i + j: O index i + j - 1
        mad_chain<4>(O + 0, a + 1, a[0]);     // a0 * (a1, a3, a5, a7) -> limbs 0..7, carry -> 8
        mad_chain<3>(O + 2, a + 2, a[1]);     // a1 * (a2, a4, a6)     -> 2..7,  carry -> 8
        mad_chain<3>(O + 4, a + 3, a[2]);     // a2 * (a3, a5, a7)     -> 4..9,  carry -> 10
        mad_chain<2>(O + 6, a + 4, a[3]);     // a3 * (a4, a6)         -> 6..9,  carry -> 10
        mad_chain<2>(O + 8, a + 5, a[4]);     // a4 * (a5, a7)         -> 8..11, carry -> 12
        mad_chain<1>(O + 10, a + 6, a[5]);    // a5 * a6               -> 10,11, carry -> 12
        mad_chain<1>(O + 12, a + 7, a[6]);    // a6 * a7               -> 12,13, carry -> 14
        // even positions i + j: E index i + j
        mad_chain<3>(E + 2, a + 2, a[0]);     // a0 * (a2, a4, a6)     -> 2..7,  carry -> 8
        mad_chain<3>(E + 4, a + 3, a[1]);     // a1 * (a3, a5, a7)     -> 4..9,  carry -> 10
        mad_chain<2>(E + 6, a + 4, a[2]);     // a2 * (a4, a6)         -> 6..9,  carry -> 10
        mad_chain<2>(E + 8, a + 5, a[3]);     // a3 * (a5, a7)         -> 8..11, carry -> 12
        mad_chain<1>(E + 10, a + 6, a[4]);    // a4 * a6               -> 10,11, carry -> 12
        mad_chain<1>(E + 12, a + 7, a[5]);    // a5 * a7               -> 12,13, carry -> 14
        // X = E + 2^32 O  (sum of the cross products, < 2^509)
        uint32_t S[16], X[16], T[16], D[16];
        S[0] = 0;
#pragma unroll
        for (int i = 1; i < 16; ++i) S[i] = O[i - 1];
        add16(X, E, S);
        // 2 X, and the diagonal a_i^2 at limbs 2i, 2i + 1
        S[0] = X[0] << 1;
#pragma unroll
        for (int i = 1; i < 16; ++i) S[i] = (X[i] << 1) | (X[i - 1] >> 31);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint64_t d = (uint64_t)a[i] * a[i];
            D[2 * i] = (uint32_t)d;
            D[2 * i + 1] = (uint32_t)(d >> 32);
        }
        add16(T, S, D);
        // Montgomery reduction of the low half
        uint32_t m[8];
        modulus(m);
        uint32_t stray = 0;
        uint32_t RE[9], RO[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) { RE[i] = T[i]; RO[i] = 0; }
        RE[8] = 0; RO[8] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t c0 = add_carry_out(RE[0], stray);
            const uint32_t mi = mul_lo(RE[0], P::INV);
            RE[8] += mad_row4(RE, m[0], m[2], m[4], m[6], mi);
            mad_row4_cin(RO, c0, m[1], m[3], m[5], m[7], mi);
            stray = RE[1];
            uint32_t nO[9];
#pragma unroll
            for (int j = 0; j < 7; ++j) nO[j] = RE[j + 2];
            nO[7] = 0; nO[8] = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) RE[j] = RO[j];
            RE[8] = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) RO[j] = nO[j];
        }
        // Q = RE + stray + 2^32 RO (<= p), result = Q + T[8..15] (< 2p)
        Fp q, r;
        uint32_t sh[8];
        sh[0] = stray;
#pragma unroll
        for (int j = 1; j < 8; ++j) sh[j] = RO[j - 1];
        add8(q.l, RE, sh);
        add8(r.l, q.l, T + 8);
        return r;
    }

    // canonical integer -> Montgomery, and back
    ZK_HD Fp to_mont() const { return (*this) * r2(); }
    ZK_HD Fp from_mont() const {   // canonical integer (fully reduced)
        Fp o = zero();
        o.l[0] = 1;
        return ((*this) * o).canon();
    }

    // a^e for a 256-bit exponent given as 8 limbs (variable time)
    ZK_HD Fp pow(const uint32_t* e) const {
        Fp r = one();
        for (int i = 255; i >= 0; --i) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1u) r = r * (*this);
        }
        return r;
    }
    ZK_HD Fp pow_u64(uint64_t e) const {
        Fp r = one();
        Fp b = *this;
        while (e) {
            if (e & 1) r = r * b;
            b = b.sqr();
            e >>= 1;
        }
        return r;
    }
    // Fermat inverse (0 -> 0)
    ZK_HD Fp inverse() const {
        uint32_t e[8];
        modulus(e);
        e[0] -= 2;  // p - 2; low limb of both moduli is >= 2
        return pow(e);
    }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace zk
