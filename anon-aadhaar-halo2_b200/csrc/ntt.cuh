// Fr number-theoretic transform kernels for sm_100a.
//
// Device replacement for halo2_proofs 0.2.0 `arithmetic::best_fft` and the
// `EvaluationDomain` wrappers `lagrange_to_coeff`, `coeff_to_extended`,
// `extended_to_coeff`, `divide_by_vanishing_poly` ([DEP] halo2_proofs/src/arithmetic.rs,
// halo2_proofs/src/poly/domain.rs @ v2023_01_20, reference Cargo.lock:469-471).
// Same contract as upstream: natural order in, natural order out,
// out[i] = sum_j a[j] * omega^(i*j).
//
// Decomposition (validated in scratch/ntt_model.py against the definition):
// n = n_1 * ... * n_P, n_p = 2^b_p, b_p <= 9.  Pass p views the buffer as
// [j_p][J][I] (|I| = n_1..n_{p-1}, |J| = n_{p+1}..n_P), transforms along j_p, multiplies
// by the inter-pass twiddle (omega^|I|)^(i_p * J) and writes [J][i_p][I].  After the last
// pass the layout is [i_P]..[i_1] = natural order, so no bit-reversal pass exists.
// One CTA owns a tile of 2^b_p rows x C columns = 2048 elements (64 KiB of shared
// memory); every thread keeps 8 elements in registers, runs radix-8/4/2 butterflies on
// them, and exchanges through shared memory between rounds.  Rows of a tile are
// C*32-byte contiguous segments in HBM; the first pass writes its whole tile as one
// contiguous 64 KiB block.
#pragma once
#include "field.cuh"

namespace zk {

#ifndef B200ZK_NTT_TILE_LOG
#define B200ZK_NTT_TILE_LOG 11
#endif
// resident CTAs per SM the pass kernel is compiled for.  Measured on B200 (scratch/r2_ntt_sweep.py): 3 (80
// registers, ~350 B of spills, 24 warps per SM) beats 2 (112 registers, no spills) at every size — 4.11 -> 3.86 ms
// at 2^24, 1.18 -> 1.11 ms for 242 transforms of 2^15 — and 1024-element tiles with 4 or 5 CTAs are slower (4.36 ms).
#ifndef B200ZK_NTT_MIN_CTAS
#define B200ZK_NTT_MIN_CTAS 3
#endif
constexpr int NTT_TILE_LOG = B200ZK_NTT_TILE_LOG;   // 2048 elements per CTA
constexpr int NTT_TILE = 1 << NTT_TILE_LOG;
constexpr int NTT_THREADS = NTT_TILE / 8;    // 256
constexpr int NTT_MAX_B = 9;

enum NttInMode : uint32_t { NTT_IN_PLAIN = 0, NTT_IN_MOD3 = 1, NTT_IN_TABLE = 2 };
enum NttOutMode : uint32_t { NTT_OUT_PLAIN = 0, NTT_OUT_MOD3 = 1 };

struct NttPassArgs {
    const Fr* in;
    Fr* out;
    uint32_t log_n;      // log2 of the whole transform
    uint32_t log_I;      // log2 |I| (product of earlier pass sizes)
    uint32_t log_cols;   // log_n - b_p : number of columns of this pass
    uint32_t last;       // 1 if this is the last pass (no inter-pass twiddle)
    const Fr* tw_tile;   // (omega^(n/n_p))^e, e < n_p
    const Fr* tw_lo;     // omega^x, x < 2^tw_h
    const Fr* tw_hi;     // omega^(y << tw_h)
    uint32_t tw_h;
    Fr w8[3];            // (omega^(n/8))^{1,2,3}
    // first pass only
    uint32_t in_mode;
    uint32_t n_in;       // rows >= n_in read as zero (zero padding)
    Fr in_tab[3];        // NTT_IN_MOD3: a[j] *= in_tab[j % 3] for j % 3 != 0
    const Fr* in_table;  // NTT_IN_TABLE: a[j] *= in_table[j & in_table_mask]
    uint32_t in_table_mask;
    // last pass only
    uint32_t out_mode;
    Fr out_tab[3];       // NTT_OUT_MOD3: out[i] *= out_tab[i % 3]
    uint32_t n_keep;     // outputs with i >= n_keep are not stored
    uint64_t in_batch_stride;   // elements between consecutive transforms of a batch
    uint64_t out_batch_stride;
    uint32_t block_offset;      // first column tile of this launch (a pass may be launched in column ranges)
    const Fr* tw_direct;        // omega^(x << log_I) for every x = i_p * J this pass can form, or null (lo / hi tables)
    // first pass of a transform sharded over several GPUs (b200zk_ntt4_first_pass_scatter_dev): `in` is this rank's
    // [2^b][2^log_cols] slab of columns col_base .. col_base + 2^log_cols of the whole transform, and output (i_p, column)
    // is stored at scatter_dest[i_p >> scatter_log_rows] + (i_p mod 2^scatter_log_rows) * scatter_pitch + scatter_col_offset
    // + (column - col_base): row i_p of the rank that owns it (a peer mapping, or a slice of a local send buffer)
    uint32_t col_base;
    uint32_t scatter_log_rows;
    uint64_t scatter_pitch, scatter_col_offset;
    Fr* scatter_dest[8];
};

// Per-(omega, log_n) twiddle tables, built on the device once and cached (ntt.cu).
struct NttTables {
    Fr omega;
    uint32_t log_n;
    int npass;
    int bits[4];
    Fr* tw_tile[NTT_MAX_B + 1];  // indexed by b: (omega^(n / 2^b))^e
    Fr* tw_lo;                   // omega^x, x < 2^tw_h
    Fr* tw_hi;                   // omega^(y << tw_h)
    uint32_t tw_h;
    Fr* tw_direct[4];            // per pass boundary p: (omega^(2^log_I_p))^x, x < n / 2^log_I_p, or null
    Fr w8[3];
    Fr* block;  // single allocation
};

#if defined(__CUDACC__)

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

template <int R> __device__ __forceinline__ constexpr int bitrev_c(int x) {
    int r = 0;
    for (int i = 0; i < R; ++i) r |= ((x >> i) & 1) << (R - 1 - i);
    return r;
}

// In-register decimation-in-frequency DFT of size 2^R on v[0..2^R): natural order in,
// slot s holds X[bitrev(s)] on return.  Root of the size-8 transform is w8[0].
template <int R> __device__ __forceinline__ void dft_regs(Fr* v, const Fr* w8) {
    if constexpr (R == 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            Fr u = v[j], t = v[j + 4];
            v[j] = u + t;
            Fr d = u - t;
            v[j + 4] = (j == 0) ? d : d * w8[j - 1];
        }
#pragma unroll
        for (int blk = 0; blk < 8; blk += 4) {
            Fr u = v[blk], t = v[blk + 2];
            v[blk] = u + t; v[blk + 2] = u - t;
            u = v[blk + 1]; t = v[blk + 3];
            v[blk + 1] = u + t; v[blk + 3] = (u - t) * w8[1];
        }
#pragma unroll
        for (int blk = 0; blk < 8; blk += 2) {
            Fr u = v[blk], t = v[blk + 1];
            v[blk] = u + t; v[blk + 1] = u - t;
        }
    } else if constexpr (R == 2) {
        Fr u = v[0], t = v[2];
        v[0] = u + t; v[2] = u - t;
        u = v[1]; t = v[3];
        v[1] = u + t; v[3] = (u - t) * w8[1];
        u = v[0]; t = v[1];
        v[0] = u + t; v[1] = u - t;
        u = v[2]; t = v[3];
        v[2] = u + t; v[3] = u - t;
    } else {
        Fr u = v[0], t = v[1];
        v[0] = u + t; v[1] = u - t;
    }
}

template <int B> struct NttRounds {
    static constexpr int NR = (B + 2) / 3;
    static constexpr int r(int q) { return (q < B / 3) ? 3 : (B % 3); }
    // log2 stride of round q's digit inside the tile position
    static constexpr int lst(int q) {
        int s = 0;
        for (int t = q + 1; t < NR; ++t) s += r(t);
        return s;
    }
    // output index digit weight: i_p = sum_q c_q << low(q)
    static constexpr int low(int q) {
        int s = 0;
        for (int t = 0; t < q; ++t) s += r(t);
        return s;
    }
};

// MODE 0: a later pass; 1: the first pass (input modifiers, contiguous tile output); 2: the first pass of a transform
// sharded over several GPUs (outputs scattered to the owners of their rows)
template <int B, int MODE, int Q>
__device__ __forceinline__ void ntt_round(const NttPassArgs& A, const Fr* __restrict__ in, Fr* __restrict__ out,
                                          uint4* S0, uint4* S1, uint32_t tid, uint32_t m0) {
    using RD = NttRounds<B>;
    constexpr bool FIRST = MODE != 0, SCATTER = MODE == 2;
    constexpr int NP = 1 << B;
    constexpr int LOGC = NTT_TILE_LOG - B;
    constexpr int C = 1 << LOGC;
    constexpr int RQ = RD::r(Q);
    constexpr int SQ = 1 << RQ;
    constexpr int LST = RD::lst(Q);
    constexpr int ST = 1 << LST;
    constexpr int G = 8 / SQ;
    constexpr bool LASTR = (Q == RD::NR - 1);
    const uint32_t ncols = 1u << A.log_cols;

    Fr v[8];
    uint32_t colv[G], basev[G], lov[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const uint32_t gamma = (uint32_t)g * NTT_THREADS + tid;  // lanes own consecutive columns
        const uint32_t col = gamma & (C - 1);
        const uint32_t other = gamma >> LOGC;
        const uint32_t lo = other & (ST - 1);
        const uint32_t hi = other >> LST;
        colv[g] = col; lov[g] = lo;
        basev[g] = (hi << (LST + RQ)) + lo;
    }
    // ---- fetch
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int a = 0; a < SQ; ++a) {
            const uint32_t pos = basev[g] + (a << LST);
            if constexpr (Q == 0) {
                const uint32_t m = m0 + colv[g];
                Fr x = Fr::zero();
                if (m < ncols) {
                    const uint32_t j = (pos << A.log_cols) + m;  // row-major [j_p][column]
                    if (!FIRST || j < A.n_in) {
                        x = ldg_fr(in + j);
                        if constexpr (FIRST) {
                            if (A.in_mode == NTT_IN_MOD3) {
                                const uint32_t r3 = j % 3u;
                                if (r3 != 0) x = x * (r3 == 1 ? A.in_tab[1] : A.in_tab[2]);
                            } else if (A.in_mode == NTT_IN_TABLE) {
                                x = x * ldg_fr(A.in_table + (j & A.in_table_mask));
                            }
                        }
                    }
                }
                v[g * SQ + a] = x;
            } else {
                const uint32_t e = (pos << LOGC) + colv[g];
                uint4 p0 = S0[e], p1 = S1[e];
                Fr x;
                x.l[0] = p0.x; x.l[1] = p0.y; x.l[2] = p0.z; x.l[3] = p0.w;
                x.l[4] = p1.x; x.l[5] = p1.y; x.l[6] = p1.z; x.l[7] = p1.w;
                v[g * SQ + a] = x;
            }
        }
    }
    // ---- butterflies
#pragma unroll
    for (int g = 0; g < G; ++g) dft_regs<RQ>(v + g * SQ, A.w8);

    if constexpr (!LASTR) {
        // twiddle by omega_{SQ*ST}^(c*lo) and write back in place at digit value c
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int a = 0; a < SQ; ++a) {
                const int c = bitrev_c<RQ>(a);
                Fr x = v[g * SQ + a];
                if (c != 0) {
                    const uint32_t e = (uint32_t)(NP >> (RQ + LST)) * (uint32_t)c * lov[g];
                    if (e != 0) x = x * ldg_fr(A.tw_tile + e);
                }
                const uint32_t pos = basev[g] + ((uint32_t)c << LST);
                const uint32_t si = (pos << LOGC) + colv[g];
                S0[si] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
                S1[si] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
            }
        }
    } else {
        if constexpr (FIRST && !SCATTER && Q > 0) __syncthreads();  // all reads of the planes done before restaging
        // ---- final: inter-pass twiddle, output modifiers, store
        Fr* stage = reinterpret_cast<Fr*>(S0);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint32_t m = m0 + colv[g];
            const bool colok = m < ncols;
            const uint32_t mg = SCATTER ? m + A.col_base : m;    // column of the whole transform
            const uint32_t Jv = mg >> A.log_I;
            const uint32_t Iv = mg & ((1u << A.log_I) - 1u);
#pragma unroll
            for (int a = 0; a < SQ; ++a) {
                const int c = bitrev_c<RQ>(a);
                const uint32_t pos = basev[g] + ((uint32_t)c << LST);
                // digit-reverse pos into the output index i_p
                uint32_t ip = 0;
#pragma unroll
                for (int q = 0; q < RD::NR; ++q) {
                    const uint32_t dq = (pos >> RD::lst(q)) & ((1u << RD::r(q)) - 1u);
                    ip |= dq << RD::low(q);
                }
                Fr x = v[g * SQ + a];
                if (colok) {
                    if (!A.last) {
                        const uint32_t ex = ip * Jv;
                        if (ex != 0) {
                            Fr t;
                            if (A.tw_direct) {
                                t = ldg_fr(A.tw_direct + ex);
                            } else {
                                const uint32_t e = ex << A.log_I;
                                t = ldg_fr(A.tw_lo + (e & ((1u << A.tw_h) - 1u)));
                                const uint32_t eh = e >> A.tw_h;
                                if (eh != 0) t = t * ldg_fr(A.tw_hi + eh);
                            }
                            x = x * t;
                        }
                    }
                    const uint32_t oi = (((Jv << B) + ip) << A.log_I) + Iv;
                    if (A.last && A.out_mode == NTT_OUT_MOD3) {
                        const uint32_t r3 = oi % 3u;
                        x = x * (r3 == 0 ? A.out_tab[0] : (r3 == 1 ? A.out_tab[1] : A.out_tab[2]));
                    }
                    if (A.last) x = x.canon();   // values leave the library fully reduced
                    if constexpr (SCATTER) {
                        // eight (2^(11 - b)) consecutive lanes hold consecutive columns of row i_p: 256-byte runs
                        Fr* q = A.scatter_dest[ip >> A.scatter_log_rows] +
                                (uint64_t)(ip & ((1u << A.scatter_log_rows) - 1u)) * A.scatter_pitch + A.scatter_col_offset + m;
                        st_fr(q, x);
                    } else if constexpr (FIRST) {
                        st_fr(stage + (colv[g] << B) + ip, x);
                    } else {
                        if (!A.last || oi < A.n_keep) st_fr(out + oi, x);
                    }
                }
            }
        }
        if constexpr (FIRST && !SCATTER) {
            // tile output is one contiguous block [m0 * NP, (m0 + C) * NP) of `out`
            __syncthreads();
            const uint32_t vcols = (ncols - m0 < (uint32_t)C) ? (ncols - m0) : (uint32_t)C;
            const uint32_t nel = vcols << B;
            const uint64_t obase = (uint64_t)m0 << B;
            uint4* dst = reinterpret_cast<uint4*>(out + obase);
            const uint4* src = reinterpret_cast<const uint4*>(stage);
            uint32_t lim = nel;
            if (A.last) {  // single-pass transform: honour n_keep
                lim = (A.n_keep > obase) ? (uint32_t)min((uint64_t)nel, (uint64_t)A.n_keep - obase) : 0u;
            }
            for (uint32_t i = tid; i < 2 * lim; i += NTT_THREADS) dst[i] = src[i];
        }
    }
}

template <int B, int MODE, int Q>
__device__ __forceinline__ void ntt_rounds_from(const NttPassArgs& A, const Fr* in, Fr* out, uint4* S0, uint4* S1,
                                                uint32_t tid, uint32_t m0) {
    ntt_round<B, MODE, Q>(A, in, out, S0, S1, tid, m0);
    if constexpr (Q + 1 < NttRounds<B>::NR) {
        __syncthreads();
        ntt_rounds_from<B, MODE, Q + 1>(A, in, out, S0, S1, tid, m0);
    }
}

template <int B, int MODE>
__global__ void __launch_bounds__(NTT_THREADS, B200ZK_NTT_MIN_CTAS) ntt_pass_kernel(const __grid_constant__ NttPassArgs A) {
    extern __shared__ uint4 ntt_smem[];
    uint4* S0 = ntt_smem;
    uint4* S1 = ntt_smem + NTT_TILE;
    constexpr int LOGC = NTT_TILE_LOG - B;
    const uint32_t m0 = (blockIdx.x + A.block_offset) << LOGC;
    const Fr* in = A.in + (uint64_t)blockIdx.y * A.in_batch_stride;
    Fr* out = A.out + (uint64_t)blockIdx.y * A.out_batch_stride;
    ntt_rounds_from<B, MODE, 0>(A, in, out, S0, S1, threadIdx.x, m0);
}

#endif  // __CUDACC__

}  // namespace zk
