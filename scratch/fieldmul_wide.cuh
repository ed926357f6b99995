// Separated-operand Montgomery multiplication for the device: one level of Karatsuba on the
// 8 x 8 limb product (48 instead of 64 IMAD.WIDE), dedicated squaring (36), and a stand-alone
// Montgomery reduction of a 512-bit value (64 IMAD.WIDE + 8 IMAD), so that sums of products can
// share one reduction.  Measured motivation: IMAD.WIDE.U32 occupies the only pipe that
// multiplies 32 x 32 -> 64 for 4 cycles per warp (32 lanes/clk/SM) while the ALU pipe that
// executes the extra additions of Karatsuba sits at ~30 % in the MSM and NTT kernels
// (profiles/r1_summary.md), so trading multiplies for adds moves work to an idle unit.
//
// Device only (inline PTX carry chains).  field.cuh keeps the CIOS form for the host
// emulation; tests/test_field_gpu.py compares the two bit for bit on edge and random values.
#pragma once
#include <stdint.h>

namespace zk {
namespace wide {

#ifdef __CUDACC__

// (a0, a1, a2, a3) += x0 * y + (x1 * y << 64); returns the carry out of a3
__device__ __forceinline__ uint32_t chain2(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3, uint32_t x0,
                                           uint32_t x1, uint32_t y) {
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "=r"(c)
        : "r"(x0), "r"(x1), "r"(y));
    return c;
}
// same when no carry can leave a3
__device__ __forceinline__ void chain2_top(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3, uint32_t x0,
                                           uint32_t x1, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
        "madc.hi.u32 %3, %5, %6, %3;"
        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3)
        : "r"(x0), "r"(x1), "r"(y));
}

// r[0..7] = a[0..3] * b[0..3].  Products whose column index is even accumulate in e[], the
// others in o[] (o[k] has weight 2^(32 (k + 1))): every partial product is one aligned 64-bit
// multiply-add, and the two accumulators are merged by one 7-limb addition at the end.
__device__ __forceinline__ void mul4x4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t e0, e1, e2, e3, e4 = 0, e5 = 0, e6, e7 = 0;
    uint32_t o0, o1, o2, o3, o4, o5 = 0, o6;
    uint64_t p;
    p = (uint64_t)a[0] * b[0]; e0 = (uint32_t)p; e1 = (uint32_t)(p >> 32);
    p = (uint64_t)a[2] * b[0]; e2 = (uint32_t)p; e3 = (uint32_t)(p >> 32);
    p = (uint64_t)a[1] * b[0]; o0 = (uint32_t)p; o1 = (uint32_t)(p >> 32);
    p = (uint64_t)a[3] * b[0]; o2 = (uint32_t)p; o3 = (uint32_t)(p >> 32);
    // row 1
    o4 = chain2(o0, o1, o2, o3, a[0], a[2], b[1]);
    chain2_top(e2, e3, e4, e5, a[1], a[3], b[1]);
    // row 2
    e6 = chain2(e2, e3, e4, e5, a[0], a[2], b[2]);
    chain2_top(o2, o3, o4, o5, a[1], a[3], b[2]);
    // row 3
    o6 = chain2(o2, o3, o4, o5, a[0], a[2], b[3]);
    chain2_top(e4, e5, e6, e7, a[1], a[3], b[3]);
    r[0] = e0;
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7), "r"(o0), "r"(o1), "r"(o2), "r"(o3),
          "r"(o4), "r"(o5), "r"(o6));
}

// r[0..7] = a[0..3]^2: the six cross products once, doubled, plus the four squares
__device__ __forceinline__ void sqr4(uint32_t* r, const uint32_t* a) {
    // cross = a0a1 (pos 1) + a0a2 (2) + a0a3 (3) + a1a2 (3) + a1a3 (4) + a2a3 (5), 7 limbs from pos 1
    uint32_t e2, e3, e4 = 0, e5 = 0;      // even columns: a0a2 (2,3), a1a3 (4,5)
    uint32_t o0, o1, o2, o3, o4, o5;      // odd columns, o[k] at position k + 1: a0a1 (1,2), a0a3 + a1a2 (3,4), a2a3 (5,6)
    uint64_t p;
    p = (uint64_t)a[0] * a[2]; e2 = (uint32_t)p; e3 = (uint32_t)(p >> 32);
    p = (uint64_t)a[0] * a[1]; o0 = (uint32_t)p; o1 = (uint32_t)(p >> 32);
    p = (uint64_t)a[0] * a[3]; o2 = (uint32_t)p; o3 = (uint32_t)(p >> 32);
    p = (uint64_t)a[2] * a[3]; o4 = (uint32_t)p; o5 = (uint32_t)(p >> 32);
    // o2,o3 += a1 a2 with the carry into (o4, o5); e4,e5 = a1 a3
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5)
        : "r"(a[1]), "r"(a[2]));
    p = (uint64_t)a[1] * a[3]; e4 = (uint32_t)p; e5 = (uint32_t)(p >> 32);
    // c[1..7] = even (positions 2..5) + odd (positions 1..6)
    uint32_t c1 = o0, c2, c3, c4, c5, c6, c7;
    asm("add.cc.u32 %0, %6, %11;\n\t"
        "addc.cc.u32 %1, %7, %12;\n\t"
        "addc.cc.u32 %2, %8, %13;\n\t"
        "addc.cc.u32 %3, %9, %14;\n\t"
        "addc.cc.u32 %4, %10, 0;\n\t"
        "addc.u32 %5, 0, 0;"
        : "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7)
        : "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(e2), "r"(e3), "r"(e4), "r"(e5));
    // double: 2 * cross < 2^256 (it is part of a^2 < 2^256)
    c7 = (c7 << 1) | (c6 >> 31); c6 = (c6 << 1) | (c5 >> 31); c5 = (c5 << 1) | (c4 >> 31);
    c4 = (c4 << 1) | (c3 >> 31); c3 = (c3 << 1) | (c2 >> 31); c2 = (c2 << 1) | (c1 >> 31); c1 <<= 1;
    // + squares at positions 0, 2, 4, 6
    uint32_t d0, d1, d2, d3, d4, d5, d6, d7;
    p = (uint64_t)a[0] * a[0]; d0 = (uint32_t)p; d1 = (uint32_t)(p >> 32);
    p = (uint64_t)a[1] * a[1]; d2 = (uint32_t)p; d3 = (uint32_t)(p >> 32);
    p = (uint64_t)a[2] * a[2]; d4 = (uint32_t)p; d5 = (uint32_t)(p >> 32);
    p = (uint64_t)a[3] * a[3]; d6 = (uint32_t)p; d7 = (uint32_t)(p >> 32);
    r[0] = d0;
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(d1), "r"(d2), "r"(d3), "r"(d4), "r"(d5), "r"(d6), "r"(d7), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
          "r"(c5), "r"(c6), "r"(c7));
}

// r[0..3] = a[0..3] + b[0..3]; returns the carry
__device__ __forceinline__ uint32_t add4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return c;
}

// z[4..8] += x[0..3] & mask   (z[8] takes the carry)
__device__ __forceinline__ void add4_masked_hi(uint32_t* z, const uint32_t* x, uint32_t mask) {
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(z[4]), "+r"(z[5]), "+r"(z[6]), "+r"(z[7]), "+r"(z[8])
        : "r"(x[0] & mask), "r"(x[1] & mask), "r"(x[2] & mask), "r"(x[3] & mask));
}

// z[0..8] -= y[0..7]   (no borrow can leave z[8]: the caller guarantees z >= y)
__device__ __forceinline__ void sub9_8(uint32_t* z, const uint32_t* y) {
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, 0;"
        : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3]), "+r"(z[4]), "+r"(z[5]), "+r"(z[6]), "+r"(z[7]), "+r"(z[8])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
}

// T[4..15] += z[0..8]  (carry rippled to the top; the total is < 2^512)
__device__ __forceinline__ void add_mid(uint32_t* T, const uint32_t* z) {
    asm("add.cc.u32 %0, %0, %12;\n\t"
        "addc.cc.u32 %1, %1, %13;\n\t"
        "addc.cc.u32 %2, %2, %14;\n\t"
        "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t"
        "addc.cc.u32 %5, %5, %17;\n\t"
        "addc.cc.u32 %6, %6, %18;\n\t"
        "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t"
        "addc.cc.u32 %9, %9, 0;\n\t"
        "addc.cc.u32 %10, %10, 0;\n\t"
        "addc.u32 %11, %11, 0;"
        : "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]),
          "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]));
}

// T[0..15] = a[0..7] * b[0..7], one level of Karatsuba:
//   a = a0 + a1 W, b = b0 + b1 W (W = 2^128);  a b = z0 + (zm - z0 - z2) W + z2 W^2,
//   zm = (a0 + a1)(b0 + b1), the two 129-bit sums handled as 4 limbs + a carry bit each.
__device__ __forceinline__ void mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    mul4x4(T, a, b);
    mul4x4(T + 8, a + 4, b + 4);
    uint32_t sa[4], sb[4], zm[9];
    const uint32_t ca = add4(sa, a, a + 4), cb = add4(sb, b, b + 4);
    mul4x4(zm, sa, sb);
    zm[8] = ca & cb;
    add4_masked_hi(zm, sb, 0u - ca);
    add4_masked_hi(zm, sa, 0u - cb);
    sub9_8(zm, T);
    sub9_8(zm, T + 8);
    add_mid(T, zm);
}

// T[0..15] = a[0..7]^2 = a0^2 + 2 a0 a1 W + a1^2 W^2
__device__ __forceinline__ void sqr_wide(uint32_t* T, const uint32_t* a) {
    sqr4(T, a);
    sqr4(T + 8, a + 4);
    uint32_t x[9];
    mul4x4(x, a, a + 4);
    x[8] = x[7] >> 31;
#pragma unroll
    for (int k = 7; k > 0; --k) x[k] = (x[k] << 1) | (x[k - 1] >> 31);
    x[0] <<= 1;
    add_mid(T, x);
}

// T += U (16 limbs; the caller guarantees the sum is < 2^512)
__device__ __forceinline__ void add_wide(uint32_t* T, const uint32_t* U) {
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.cc.u32 %7, %7, %15;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7])
        : "r"(U[0]), "r"(U[1]), "r"(U[2]), "r"(U[3]), "r"(U[4]), "r"(U[5]), "r"(U[6]), "r"(U[7]));
    asm("addc.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(U[8]), "r"(U[9]), "r"(U[10]), "r"(U[11]), "r"(U[12]), "r"(U[13]), "r"(U[14]), "r"(U[15]));
}

// One reduction row on the odd accumulator: lo += stray; m = lo * inv; O += (x0, x1, x2, x3) * m
// with the carry of the first addition entering the chain.  No carry leaves limb 7 (see
// mont_reduce).
__device__ __forceinline__ uint32_t reduce_row_odd(uint32_t& lo, uint32_t stray, uint32_t inv, uint32_t* acc, uint32_t x0,
                                                   uint32_t x1, uint32_t x2, uint32_t x3) {
    uint32_t m;
    asm("add.cc.u32 %8, %8, %10;\n\t"
        "mul.lo.u32 %9, %8, %11;\n\t"
        "madc.lo.cc.u32 %0, %12, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %12, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %13, %9, %2;\n\t"
        "madc.hi.cc.u32 %3, %13, %9, %3;\n\t"
        "madc.lo.cc.u32 %4, %14, %9, %4;\n\t"
        "madc.hi.cc.u32 %5, %14, %9, %5;\n\t"
        "madc.lo.cc.u32 %6, %15, %9, %6;\n\t"
        "madc.hi.u32 %7, %15, %9, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(lo), "=&r"(m)
        : "r"(stray), "r"(inv), "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    return m;
}

// acc(8 limbs) += (x0, x1, x2, x3) * y at even positions; returns the carry out of limb 7
__device__ __forceinline__ uint32_t reduce_row_even(uint32_t* acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                                                    uint32_t y) {
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
}

// r = T / 2^256 mod p for T < 2^256 * p (then r < 2p):  r = (T + M p) / 2^256 with
// M = sum m_i 2^(32 i) chosen row by row so the low limb cancels.  The running window
// W = E + stray + 2^32 O (E nine limbs, O eight) starts as T[0..7]; row i adds m_i p (even
// limbs of p into E, odd limbs into O), drops the now-zero low limb (register renaming, as in
// field.cuh) and brings T[i + 8] in at the top.  After row i, W < 2^256 + p, hence O < 2^225:
// no carry leaves O, and E needs its ninth limb.
template <class P>
__device__ __forceinline__ void mont_reduce(uint32_t* r, const uint32_t* T) {
    uint32_t E[9], O[9];
#pragma unroll
    for (int j = 0; j < 8; ++j) { E[j] = T[j]; O[j] = 0; }
    E[8] = 0; O[8] = 0;
    uint32_t stray = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t m = reduce_row_odd(E[0], stray, P::INV, O, P::P1, P::P3, P::P5, P::P7);
        E[8] += reduce_row_even(E, P::P0, P::P2, P::P4, P::P6, m);
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
        // T[i + 8] enters at limb 7 of the shifted window
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.u32 %1, %1, 0;"
            : "+r"(E[7]), "+r"(E[8])
            : "r"(T[i + 8]));
    }
    // r = E + stray + 2^32 O
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(stray),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
}

#endif  // __CUDACC__

}  // namespace wide
}  // namespace zk
