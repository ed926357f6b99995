// Column arithmetic around the commitments (SURVEY.md section 8 f, ranks 1 and 2): the loops
// `create_proof` runs immediately before and after the MSM / NTT hot path.  With those on the
// GPU, these would otherwise force every column through host memory.
//
// Upstream items replaced ([DEP] halo2_proofs 0.2.0 @ v2023_01_20, reference Cargo.lock:469-471;
// ff 0.12 `BatchInvert`, reference Cargo.lock:339-341):
//   ff::BatchInvert::batch_invert                      -> b200zk_batch_invert(_dev)
//   plonk/permutation/prover.rs  Argument::commit      -> b200zk_permutation_product_dev
//   plonk/lookup/prover.rs       Permuted::commit_product -> b200zk_lookup_product_dev
//   arithmetic.rs                eval_polynomial       -> b200zk_eval_polynomial_dev
//   arithmetic.rs                kate_division         -> b200zk_kate_division_dev
// All results are the unique field elements the CPU code computes (no floating point, no
// ordering freedom), so parity is bit-exact by construction of canonical outputs.
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ntt.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace zk {

constexpr int PL = 8;            // elements per thread
constexpr int PT = 256;          // threads per block
constexpr int PB = PL * PT;      // elements per block

// ------------------------------------------------------------------ batch inversion
// Montgomery's trick, hierarchically: groups of 8 (strided inside a 2048-element tile so that
// loads coalesce) are multiplied together, the group products are inverted recursively, and
// each group back-substitutes.  ~4.6 multiplications per element instead of ~380.  Zeros are
// skipped and stay zero, as in ff::BatchInvert.
static __global__ void __launch_bounds__(PT) inv_group_product_kernel(const Fr* __restrict__ a, size_t n,
                                                                      Fr* __restrict__ prod) {
    const size_t base = (size_t)blockIdx.x * PB + threadIdx.x;
    Fr p = Fr::one();
#pragma unroll
    for (int j = 0; j < PL; ++j) {
        const size_t idx = base + (size_t)j * PT;
        if (idx < n) {
            const Fr x = ld_fr(a + idx);
            if (!x.is_zero()) p = p * x;
        }
    }
    st_fr(prod + (size_t)blockIdx.x * PT + threadIdx.x, p);
}

static __global__ void __launch_bounds__(PT) inv_group_apply_kernel(Fr* __restrict__ a, size_t n,
                                                                    const Fr* __restrict__ prod_inv) {
    const size_t base = (size_t)blockIdx.x * PB + threadIdx.x;
    Fr x[PL], pre[PL];
    bool live[PL];
    Fr run = Fr::one();
#pragma unroll
    for (int j = 0; j < PL; ++j) {
        const size_t idx = base + (size_t)j * PT;
        live[j] = false;
        if (idx < n) {
            x[j] = ld_fr(a + idx);
            live[j] = !x[j].is_zero();
        }
        pre[j] = run;                          // product of the live elements before j
        if (live[j]) run = run * x[j];
    }
    Fr inv = ld_fr(prod_inv + (size_t)blockIdx.x * PT + threadIdx.x);
#pragma unroll
    for (int j = PL - 1; j >= 0; --j) {
        if (live[j]) {
            const size_t idx = base + (size_t)j * PT;
            st_fr(a + idx, (inv * pre[j]).canon());
            inv = inv * x[j];
        }
    }
}

static __global__ void inv_direct_kernel(Fr* a, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr x = ld_fr(a + i);
    if (!x.is_zero()) st_fr(a + i, x.inverse().canon());
    else st_fr(a + i, Fr::zero());
}

static size_t batch_invert_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > PB) {
        const size_t groups = (n + PB - 1) / PB * PT;
        total += groups;
        n = groups;
    }
    return total + 1;
}

static void batch_invert_run(Fr* a, size_t n, Fr* scratch, cudaStream_t s) {
    if (n == 0) return;
    if (n <= (size_t)PB) {
        inv_direct_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a, n);
        ZK_LAUNCH_CHECK();
        return;
    }
    const size_t nblocks = (n + PB - 1) / PB, groups = nblocks * PT;
    inv_group_product_kernel<<<(unsigned)nblocks, PT, 0, s>>>(a, n, scratch);
    ZK_LAUNCH_CHECK();
    batch_invert_run(scratch, groups, scratch + groups, s);
    inv_group_apply_kernel<<<(unsigned)nblocks, PT, 0, s>>>(a, n, scratch);
    ZK_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ prefix products
// out[c][0] = first[c] (or 1), out[c][i] = out[c][i-1] * in[c][i-1]: the
// `z.push(z[row - 1] * modified_values[row - 1])` loop of the permutation argument and the
// `scan` of the lookup argument.  Three kernels: per-thread products + block scan, a serial
// scan over the (few) block totals of each column, and the application.
__device__ __forceinline__ void sh_store(uint4* sh, int i, const Fr& v) {
    sh[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sh[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr sh_load(const uint4* sh, int i) {
    const uint4 a = sh[2 * i], b = sh[2 * i + 1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

static __global__ void __launch_bounds__(PT) pp_block_kernel(const Fr* __restrict__ in, size_t in_stride, size_t n,
                                                             Fr* __restrict__ thread_prefix,
                                                             Fr* __restrict__ block_total, uint32_t nblocks) {
    __shared__ uint4 sh[2 * PT];
    const uint32_t c = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
    const Fr* col = in + (size_t)c * in_stride;
    const size_t start = (size_t)b * PB + (size_t)t * PL;
    Fr p = Fr::one();
#pragma unroll
    for (int j = 0; j < PL; ++j)
        if (start + j < n) p = p * ld_fr(col + start + j);
    sh_store(sh, t, p);
    __syncthreads();
    for (int d = 1; d < PT; d <<= 1) {          // inclusive Hillis-Steele scan
        Fr v = sh_load(sh, t);
        if ((int)t >= d) v = sh_load(sh, t - d) * v;
        __syncthreads();
        sh_store(sh, t, v);
        __syncthreads();
    }
    const Fr excl = t ? sh_load(sh, t - 1) : Fr::one();
    st_fr(thread_prefix + ((size_t)c * nblocks + b) * PT + t, excl);
    if (t == PT - 1) st_fr(block_total + (size_t)c * nblocks + b, sh_load(sh, t));
}

static __global__ void pp_scan_blocks_kernel(Fr* block_total, uint32_t nblocks, uint32_t count, const Fr* first) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    Fr run = first ? ld_fr(first + c) : Fr::one();
    for (uint32_t b = 0; b < nblocks; ++b) {
        const Fr tot = ld_fr(block_total + (size_t)c * nblocks + b);
        st_fr(block_total + (size_t)c * nblocks + b, run);
        run = run * tot;
    }
}

static __global__ void __launch_bounds__(PT) pp_apply_kernel(const Fr* in, size_t in_stride, Fr* out, size_t out_stride, size_t n,
                                                             const Fr* __restrict__ thread_prefix,
                                                             const Fr* __restrict__ block_prefix, uint32_t nblocks) {
    const uint32_t c = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
    const Fr* col = in + (size_t)c * in_stride;
    Fr* dst = out + (size_t)c * out_stride;
    const size_t start = (size_t)b * PB + (size_t)t * PL;
    if (start >= n) return;
    Fr x[PL];
#pragma unroll
    for (int j = 0; j < PL; ++j)
        if (start + j < n) x[j] = ld_fr(col + start + j);      // read before writing: out may alias in
    Fr run = ld_fr(block_prefix + (size_t)c * nblocks + b) * ld_fr(thread_prefix + ((size_t)c * nblocks + b) * PT + t);
#pragma unroll
    for (int j = 0; j < PL; ++j) {
        if (start + j < n) {
            st_fr(dst + start + j, run.canon());
            run = run * x[j];
        }
    }
}

// first: device array of `count` elements or null (all ones)
static void prefix_product_run(Context& c, const Fr* in, size_t in_stride, Fr* out, size_t out_stride, size_t count,
                               size_t n, const Fr* first, cudaStream_t s) {
    if (count == 0 || n == 0) return;
    const uint32_t nblocks = (uint32_t)((n + PB - 1) / PB);
    Fr* scratch = (Fr*)c.scratch(s).poly_scan.get((count * nblocks * (PT + 1) + 1) * sizeof(Fr));
    Fr* thread_prefix = scratch;
    Fr* block_total = scratch + count * nblocks * PT;
    pp_block_kernel<<<dim3(nblocks, (unsigned)count), PT, 0, s>>>(in, in_stride, n, thread_prefix, block_total, nblocks);
    ZK_LAUNCH_CHECK();
    pp_scan_blocks_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(block_total, nblocks, (uint32_t)count, first);
    ZK_LAUNCH_CHECK();
    pp_apply_kernel<<<dim3(nblocks, (unsigned)count), PT, 0, s>>>(in, in_stride, out, out_stride, n, thread_prefix,
                                                                 block_total, nblocks);
    ZK_LAUNCH_CHECK();
}

// ------------------------------------------------------------- permutation argument
struct PermProdArgs {
    const Fr* const* values;     // n_cols column pointers (Lagrange basis, n rows)
    const Fr* const* sigma;      // n_cols permutation polynomials in Lagrange basis (pkey.permutations)
    const Fr* delta_beta;        // n_cols: delta^j * beta
    Fr* work;                    // n_sets x n
    const Fr* tw_lo;
    const Fr* tw_hi;
    uint32_t tw_h;
    uint32_t n_cols, chunk_len, log_n;
    Fr beta, gamma;
};

// denominators: prod_col (beta * sigma_col[i] + gamma + v_col[i])
static __global__ void perm_denominator_kernel(const __grid_constant__ PermProdArgs A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = 1u << A.log_n, set = blockIdx.y;
    if (i >= n) return;
    const uint32_t lo = set * A.chunk_len, hi = min(lo + A.chunk_len, A.n_cols);
    Fr d = Fr::one();
    for (uint32_t col = lo; col < hi; ++col)
        d = d * (A.beta * ldg_fr(A.sigma[col] + i) + A.gamma + ldg_fr(A.values[col] + i));
    st_fr(A.work + (size_t)set * n + i, d);
}

// modified_values = inverted denominator * prod_col (delta^j * beta * omega^i + gamma + v_col[i])
static __global__ void perm_numerator_kernel(const __grid_constant__ PermProdArgs A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = 1u << A.log_n, set = blockIdx.y;
    if (i >= n) return;
    const uint32_t lo = set * A.chunk_len, hi = min(lo + A.chunk_len, A.n_cols);
    Fr w = ldg_fr(A.tw_lo + (i & ((1u << A.tw_h) - 1u)));
    const uint32_t ih = i >> A.tw_h;
    if (ih != 0) w = w * ldg_fr(A.tw_hi + ih);                // omega^i
    Fr m = ld_fr(A.work + (size_t)set * n + i);
    for (uint32_t col = lo; col < hi; ++col)
        m = m * (ldg_fr(A.delta_beta + col) * w + A.gamma + ldg_fr(A.values[col] + i));
    st_fr(A.work + (size_t)set * n + i, m);
}

// chain[s] = z_s[0]: chain[0] = 1, chain[s] = chain[s-1] * P_{s-1}[last_row]   (P_s = raw prefix products)
static __global__ void perm_chain_kernel(const Fr* z, uint32_t n, uint32_t n_sets, uint32_t last_row, Fr* chain) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr run = Fr::one();
    for (uint32_t s = 0; s < n_sets; ++s) {
        st_fr(chain + s, run);
        run = run * ld_fr(z + (size_t)s * n + last_row);
    }
}

// z_s[i] *= chain[s]; rows >= n - n_blind receive the caller's blinding scalars when given
static __global__ void perm_scale_kernel(Fr* z, uint32_t n, const Fr* chain, const Fr* blinds, uint32_t n_blind) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (i >= n) return;
    Fr* p = z + (size_t)s * n + i;
    if (blinds && i >= n - n_blind) {
        st_fr(p, ldg_fr(blinds + (size_t)s * n_blind + (i - (n - n_blind))).canon());
        return;
    }
    st_fr(p, (ld_fr(p) * ldg_fr(chain + s)).canon());
}

// --------------------------------------------------------------------- lookup argument
struct LookupProdArgs {
    const Fr* const* compressed_input;
    const Fr* const* compressed_table;
    const Fr* const* permuted_input;
    const Fr* const* permuted_table;
    Fr* work;            // count x n
    uint32_t log_n;
    Fr beta, gamma;
};

static __global__ void lookup_denominator_kernel(const __grid_constant__ LookupProdArgs A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
    const uint32_t n = 1u << A.log_n;
    if (i >= n) return;
    const Fr d = (A.beta + ldg_fr(A.permuted_input[l] + i)) * (A.gamma + ldg_fr(A.permuted_table[l] + i));
    st_fr(A.work + (size_t)l * n + i, d);
}

static __global__ void lookup_numerator_kernel(const __grid_constant__ LookupProdArgs A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
    const uint32_t n = 1u << A.log_n;
    if (i >= n) return;
    Fr m = ld_fr(A.work + (size_t)l * n + i);
    m = m * (ldg_fr(A.compressed_input[l] + i) + A.beta);
    m = m * (ldg_fr(A.compressed_table[l] + i) + A.gamma);
    st_fr(A.work + (size_t)l * n + i, m);
}

static __global__ void blind_rows_kernel(Fr* z, uint32_t n, const Fr* blinds, uint32_t n_blind) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (j >= n_blind) return;
    st_fr(z + (size_t)s * n + (n - n_blind) + j, ldg_fr(blinds + (size_t)s * n_blind + j).canon());
}

// --------------------------------------------------------------------- eval_polynomial
// sum_i a[i] x^i.  Level kernel: every thread Horner-evaluates 8 consecutive coefficients, the
// block folds its 256 partials with the powers x^8, x^16, ... (binary tree in shared memory)
// and writes one value; the next level evaluates those at x^2048.  blockIdx.y = polynomial.
static __global__ void __launch_bounds__(PT) eval_level_kernel(const Fr* __restrict__ in, size_t in_stride, size_t n,
                                                               const Fr* __restrict__ points, Fr* __restrict__ out,
                                                               size_t out_stride) {
    __shared__ uint4 sh[2 * PT];
    const uint32_t c = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
    const Fr* col = in + (size_t)c * in_stride;
    const Fr x = ldg_fr(points + c);
    const size_t start = (size_t)b * PB + (size_t)t * PL;
    Fr acc = Fr::zero();
#pragma unroll
    for (int j = PL - 1; j >= 0; --j)
        if (start + j < n) acc = acc * x + ld_fr(col + start + j);
    sh_store(sh, t, acc);
    Fr y = x;
#pragma unroll
    for (int j = 1; j < PL; j <<= 1) y = y * y;             // x^PL
    __syncthreads();
    for (int d = 1; d < PT; d <<= 1) {
        if ((t & (2 * d - 1)) == 0) {
            const Fr v = sh_load(sh, t) + sh_load(sh, t + d) * y;
            sh_store(sh, t, v);
        }
        y = y * y;
        __syncthreads();
    }
    if (t == 0) st_fr(out + (size_t)c * out_stride + b, sh_load(sh, 0).canon());
}

static __global__ void pow_points_kernel(const Fr* in, Fr* out, uint32_t count, uint32_t log_e) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    Fr y = ld_fr(in + c);
    for (uint32_t j = 0; j < log_e; ++j) y = y * y;
    st_fr(out + c, y);
}

// ----------------------------------------------------------------------- kate_division
// q = a / (X - b) for a of n coefficients: q[n-2] = a[n-1], q[i-1] = a[i] + b * q[i]; i.e.
// q[j] = sum_{i > j} a[i] b^(i-j-1).  Suffix scan of the affine maps v -> v * b + a[i]:
// block sums (the eval kernel at x = b over 2048-coefficient tiles), a serial pass over the
// tiles for the carry entering each tile from the right, then the tile-local scan.
static __global__ void kate_carry_kernel(const Fr* tile_sums, uint32_t ntiles, Fr b_pow_tile, Fr* carry_in) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // carry_in[t] = sum_{i >= (t+1) * PB} a[i] b^(i - (t+1) PB)
    Fr run = Fr::zero();
    for (int t = (int)ntiles - 1; t >= 0; --t) {
        st_fr(carry_in + t, run);
        run = run * b_pow_tile + ld_fr(tile_sums + t);
    }
}

static __global__ void __launch_bounds__(PT) kate_tile_kernel(const Fr* __restrict__ a, size_t n, Fr b,
                                                              const Fr* __restrict__ carry_in, Fr* __restrict__ q) {
    __shared__ uint4 sh[2 * PT];
    const uint32_t tile = blockIdx.x, t = threadIdx.x;
    const size_t start = (size_t)tile * PB + (size_t)t * PL;
    Fr x[PL];
    Fr s = Fr::zero();                          // sum_j x[j] b^j over this thread's chunk
#pragma unroll
    for (int j = PL - 1; j >= 0; --j) {
        x[j] = (start + j < n) ? ld_fr(a + start + j) : Fr::zero();
        s = s * b + x[j];
    }
    sh_store(sh, t, s);
    Fr y = b;
#pragma unroll
    for (int j = 1; j < PL; j <<= 1) y = y * y;             // b^PL
    __syncthreads();
    // suffix scan: S[t] = sum_{u >= t} s[u] * (b^PL)^(u - t)
    for (int d = 1; d < PT; d <<= 1) {
        Fr v = sh_load(sh, t);
        if ((int)t + d < PT) v = v + sh_load(sh, t + d) * y;
        __syncthreads();
        sh_store(sh, t, v);
        y = y * y;
        __syncthreads();
    }
    // y is now b^(PL * PT) = b^PB; carry entering this thread from the right:
    //   C = S[t + 1] + (b^PL)^(PT - 1 - t) * carry_in[tile]
    Fr c_in = ldg_fr(carry_in + tile);
    Fr carry = (t + 1 < PT) ? sh_load(sh, t + 1) : Fr::zero();
    if (!c_in.is_zero()) {
        Fr bp = b;
#pragma unroll
        for (int j = 1; j < PL; j <<= 1) bp = bp * bp;      // b^PL
        carry = carry + c_in * bp.pow_u64((uint64_t)(PT - 1 - t));
    }
    // q[i - 1] = a[i] + b * q[i], walking down from the top of the chunk; `carry` = q[start + PL - 1]
    Fr run = carry;
#pragma unroll
    for (int j = PL - 1; j >= 0; --j) {
        const size_t i = start + j;                 // q[i] = run after adding the terms above i
        if (i + 1 < n) st_fr(q + i, run.canon());
        run = run * b + x[j];
    }
}

// -------------------------------------------------------------- linear combinations
// out[i] = sum_j coeff[j] * poly_j[i]: the `acc * y + poly` / `poly * scalar` chains of the
// multi-open provers (halo2_proofs/src/poly/kzg/multiopen/{shplonk,gwc}/prover.rs) collapsed
// into one pass that reads every column once (HBM-bound: count x 32 B per output element).
static __global__ void linear_combination_kernel(const Fr* const* __restrict__ polys, const Fr* __restrict__ coeffs,
                                                 uint32_t count, size_t n, Fr* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr acc = Fr::zero();
    for (uint32_t j = 0; j < count; ++j) acc = acc + ldg_fr(coeffs + j) * ldg_fr(polys[j] + i);
    st_fr(out + i, acc.canon());
}

}  // namespace zk

using namespace zk;

static cudaStream_t pick_stream(void* stream) { return stream ? (cudaStream_t)stream : ctx().stream; }

extern "C" {

int b200zk_batch_invert_dev(void* d_a, size_t n, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_a || n == 0, "null argument");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        Fr* scratch = (Fr*)c.scratch(s).poly_work.get(batch_invert_scratch_elems(n) * sizeof(Fr));
        batch_invert_run((Fr*)d_a, n, scratch, s);
    });
}

int b200zk_batch_invert(uint64_t* a, size_t n) {
    return guarded([&] {
        ZK_REQUIRE(a || n == 0, "null argument");
        if (n == 0) return;
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        Fr* d = (Fr*)c.scratch(s).ntt_io.get(n * sizeof(Fr));
        Fr* scratch = (Fr*)c.scratch(s).poly_work.get(batch_invert_scratch_elems(n) * sizeof(Fr));
        ZK_CUDA(cudaMemcpyAsync(d, a, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
        batch_invert_run(d, n, scratch, s);
        ZK_CUDA(cudaMemcpyAsync(a, d, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_prefix_product_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                              size_t n, const uint64_t* first_or_null, void* stream) {
    return guarded([&] {
        ZK_REQUIRE((d_in && d_out) || n == 0 || count == 0, "null argument");
        ZK_REQUIRE(count <= 65535, "batch count exceeds 65535");
        ZK_REQUIRE(count <= 1 || (in_stride >= n && out_stride >= n), "batch stride smaller than the column");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        Fr* first = nullptr;
        if (first_or_null && count) {
            first = (Fr*)c.scratch(s).poly_small.get(count * sizeof(Fr));
            c.scratch(s).staging.copy(first, first_or_null, count * sizeof(Fr), s);
        }
        prefix_product_run(c, (const Fr*)d_in, in_stride, (Fr*)d_out, out_stride, count, n, first, s);
        if (first) ZK_CUDA(cudaStreamSynchronize(s));   // the host array may go away after return
    });
}

int b200zk_permutation_product_dev(const void* const* d_values, const void* const* d_sigma, uint32_t n_cols,
                                   uint32_t chunk_len, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4],
                                   const uint64_t omega[4], const uint64_t delta[4], uint32_t blinding_factors,
                                   const uint64_t* blinds_or_null, void* d_z, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_values && d_sigma && beta && gamma && omega && delta && d_z, "null argument");
        ZK_REQUIRE(n_cols >= 1 && chunk_len >= 1 && k >= 1 && k <= 28, "bad permutation shape");
        const uint32_t n = 1u << k;
        ZK_REQUIRE(blinding_factors + 1 < n, "too many blinding factors");
        const uint32_t n_sets = (n_cols + chunk_len - 1) / chunk_len;
        ZK_REQUIRE(n_sets <= 65535, "too many permutation sets");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        const Fr b = fr_from_limbs(beta), g = fr_from_limbs(gamma), w = fr_from_limbs(omega), dl = fr_from_limbs(delta);
        NttTables* t = ntt_get_tables(c, w, k, s);
        // small device arrays: pointers, delta^j * beta, chain, blinds
        std::vector<Fr> db(n_cols);
        Fr cur = b;                                   // delta^0 * beta
        for (uint32_t j = 0; j < n_cols; ++j) { db[j] = cur; cur = cur * dl; }
        const size_t ptr_bytes = (size_t)n_cols * sizeof(void*);
        const size_t n_blind_elems = blinds_or_null ? (size_t)n_sets * blinding_factors : 0;
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
        const size_t o_v = carve(ptr_bytes), o_s = carve(ptr_bytes), o_db = carve(n_cols * sizeof(Fr));
        const size_t o_chain = carve(n_sets * sizeof(Fr)), o_bl = carve(std::max<size_t>(n_blind_elems, 1) * sizeof(Fr));
        char* small = (char*)c.scratch(s).poly_small.get(off);
        for (uint32_t j = 0; j < n_cols; ++j) ZK_REQUIRE(d_values[j] && d_sigma[j], "null column pointer");
        c.scratch(s).staging.copy(small + o_v, d_values, ptr_bytes, s);
        c.scratch(s).staging.copy(small + o_s, d_sigma, ptr_bytes, s);
        c.scratch(s).staging.copy(small + o_db, db.data(), n_cols * sizeof(Fr), s);
        if (n_blind_elems)
            c.scratch(s).staging.copy(small + o_bl, blinds_or_null, n_blind_elems * sizeof(Fr), s);
        PermProdArgs A;
        A.values = (const Fr* const*)(small + o_v);
        A.sigma = (const Fr* const*)(small + o_s);
        A.delta_beta = (const Fr*)(small + o_db);
        A.work = (Fr*)c.scratch(s).poly_cols.get((size_t)n_sets * n * sizeof(Fr));
        A.tw_lo = t->tw_lo; A.tw_hi = t->tw_hi; A.tw_h = t->tw_h;
        A.n_cols = n_cols; A.chunk_len = chunk_len; A.log_n = k;
        A.beta = b; A.gamma = g;
        const dim3 grid((n + 255) / 256, n_sets);
        perm_denominator_kernel<<<grid, 256, 0, s>>>(A);
        ZK_LAUNCH_CHECK();
        Fr* scratch = (Fr*)c.scratch(s).poly_work.get(batch_invert_scratch_elems((size_t)n_sets * n) * sizeof(Fr));
        batch_invert_run(A.work, (size_t)n_sets * n, scratch, s);
        perm_numerator_kernel<<<grid, 256, 0, s>>>(A);
        ZK_LAUNCH_CHECK();
        Fr* z = (Fr*)d_z;
        prefix_product_run(c, A.work, n, z, n, n_sets, n, nullptr, s);
        Fr* chain = (Fr*)(small + o_chain);
        perm_chain_kernel<<<1, 32, 0, s>>>(z, n, n_sets, n - (blinding_factors + 1), chain);
        ZK_LAUNCH_CHECK();
        perm_scale_kernel<<<grid, 256, 0, s>>>(z, n, chain, n_blind_elems ? (const Fr*)(small + o_bl) : nullptr,
                                               blinding_factors);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(s));   // host staging vectors go out of scope
    });
}

int b200zk_lookup_product_dev(const void* const* d_compressed_input, const void* const* d_compressed_table,
                              const void* const* d_permuted_input, const void* const* d_permuted_table,
                              uint32_t count, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4],
                              uint32_t blinding_factors, const uint64_t* blinds_or_null, void* d_z, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_compressed_input && d_compressed_table && d_permuted_input && d_permuted_table && beta && gamma &&
                       d_z, "null argument");
        ZK_REQUIRE(count >= 1 && count <= 65535 && k >= 1 && k <= 28, "bad lookup shape");
        const uint32_t n = 1u << k;
        ZK_REQUIRE(blinding_factors + 1 < n, "too many blinding factors");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        const size_t ptr_bytes = (size_t)count * sizeof(void*);
        const size_t n_blind_elems = blinds_or_null ? (size_t)count * blinding_factors : 0;
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
        const size_t o_p[4] = {carve(ptr_bytes), carve(ptr_bytes), carve(ptr_bytes), carve(ptr_bytes)};
        const size_t o_bl = carve(std::max<size_t>(n_blind_elems, 1) * sizeof(Fr));
        char* small = (char*)c.scratch(s).poly_small.get(off);
        const void* const* srcs[4] = {d_compressed_input, d_compressed_table, d_permuted_input, d_permuted_table};
        for (int a = 0; a < 4; ++a) {
            for (uint32_t j = 0; j < count; ++j) ZK_REQUIRE(srcs[a][j], "null column pointer");
            c.scratch(s).staging.copy(small + o_p[a], srcs[a], ptr_bytes, s);
        }
        if (n_blind_elems)
            c.scratch(s).staging.copy(small + o_bl, blinds_or_null, n_blind_elems * sizeof(Fr), s);
        LookupProdArgs A;
        A.compressed_input = (const Fr* const*)(small + o_p[0]);
        A.compressed_table = (const Fr* const*)(small + o_p[1]);
        A.permuted_input = (const Fr* const*)(small + o_p[2]);
        A.permuted_table = (const Fr* const*)(small + o_p[3]);
        A.work = (Fr*)c.scratch(s).poly_cols.get((size_t)count * n * sizeof(Fr));
        A.log_n = k;
        A.beta = fr_from_limbs(beta); A.gamma = fr_from_limbs(gamma);
        const dim3 grid((n + 255) / 256, count);
        lookup_denominator_kernel<<<grid, 256, 0, s>>>(A);
        ZK_LAUNCH_CHECK();
        Fr* scratch = (Fr*)c.scratch(s).poly_work.get(batch_invert_scratch_elems((size_t)count * n) * sizeof(Fr));
        batch_invert_run(A.work, (size_t)count * n, scratch, s);
        lookup_numerator_kernel<<<grid, 256, 0, s>>>(A);
        ZK_LAUNCH_CHECK();
        prefix_product_run(c, A.work, n, (Fr*)d_z, n, count, n, nullptr, s);
        if (n_blind_elems && blinding_factors) {
            blind_rows_kernel<<<dim3((blinding_factors + 63) / 64, count), 64, 0, s>>>((Fr*)d_z, n, (const Fr*)(small + o_bl),
                                                                                      blinding_factors);
            ZK_LAUNCH_CHECK();
        }
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_eval_polynomial_dev(const void* d_polys, size_t stride, size_t count, size_t n, const uint64_t* points,
                               uint64_t* out, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(out && (count == 0 || (points && (d_polys || n == 0))), "null argument");
        ZK_REQUIRE(count <= 65535, "batch count exceeds 65535");
        ZK_REQUIRE(count <= 1 || stride >= n, "batch stride smaller than the polynomial");
        if (count == 0) return;
        if (n == 0) {
            memset(out, 0, count * sizeof(Fr));
            return;
        }
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        // level sizes
        std::vector<size_t> sizes{n};
        do sizes.push_back((sizes.back() + PB - 1) / PB); while (sizes.back() > 1);
        size_t partial_elems = 0;
        for (size_t i = 1; i < sizes.size(); ++i) partial_elems += sizes[i] * count;
        Fr* work = (Fr*)c.scratch(s).poly_work.get((partial_elems + 2 * count + 1) * sizeof(Fr));
        Fr* pts_a = work;
        Fr* pts_b = work + count;
        Fr* partial = work + 2 * count;
        c.scratch(s).staging.copy(pts_a, points, count * sizeof(Fr), s);
        const Fr* src = (const Fr*)d_polys;
        size_t src_stride = stride;
        uint32_t log_tile = 0;
        while ((1u << log_tile) < (uint32_t)PB) ++log_tile;
        for (size_t lvl = 0; lvl + 1 < sizes.size(); ++lvl) {
            const size_t m = sizes[lvl], blocks = sizes[lvl + 1];
            eval_level_kernel<<<dim3((unsigned)blocks, (unsigned)count), PT, 0, s>>>(src, src_stride, m, pts_a, partial, blocks);
            ZK_LAUNCH_CHECK();
            src = partial;
            src_stride = blocks;
            partial += blocks * count;
            if (lvl + 2 < sizes.size()) {
                pow_points_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(pts_a, pts_b, (uint32_t)count, log_tile);
                ZK_LAUNCH_CHECK();
                std::swap(pts_a, pts_b);
            }
        }
        ZK_CUDA(cudaMemcpyAsync(out, src, count * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_kate_division_dev(const void* d_a, size_t n, const uint64_t b[4], void* d_q, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_a && b && (d_q || n <= 1), "null argument");
        ZK_REQUIRE(n >= 1, "kate_division needs at least one coefficient");
        if (n == 1) return;                       // quotient of a constant is empty
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        const Fr bb = fr_from_limbs(b);
        const uint32_t ntiles = (uint32_t)((n + PB - 1) / PB);
        Fr* work = (Fr*)c.scratch(s).poly_work.get((2 * (size_t)ntiles + 2) * sizeof(Fr));
        Fr* tile_sums = work;
        Fr* carry_in = work + ntiles;
        Fr* pt = work + 2 * (size_t)ntiles;
        c.scratch(s).staging.copy(pt, &bb, sizeof(Fr), s);
        eval_level_kernel<<<dim3(ntiles, 1), PT, 0, s>>>((const Fr*)d_a, n, n, pt, tile_sums, ntiles);
        ZK_LAUNCH_CHECK();
        kate_carry_kernel<<<1, 32, 0, s>>>(tile_sums, ntiles, bb.pow_u64((uint64_t)PB), carry_in);
        ZK_LAUNCH_CHECK();
        kate_tile_kernel<<<ntiles, PT, 0, s>>>((const Fr*)d_a, n, bb, carry_in, (Fr*)d_q);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(s));        // `bb` staging
    });
}

int b200zk_linear_combination_dev(const void* const* d_polys, const uint64_t* coeffs, uint32_t count, size_t n,
                                  void* d_out, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_out || n == 0, "null argument");
        ZK_REQUIRE(count == 0 || (d_polys && coeffs), "null argument");
        if (n == 0) return;
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = pick_stream(stream);
        const size_t ptr_bytes = (size_t)count * sizeof(void*);
        const size_t o_c = (ptr_bytes + 255) / 256 * 256;
        char* small = (char*)c.scratch(s).poly_small.get(o_c + (size_t)count * sizeof(Fr) + 256);
        for (uint32_t j = 0; j < count; ++j) ZK_REQUIRE(d_polys[j], "null column pointer");
        if (count) {
            c.scratch(s).staging.copy(small, d_polys, ptr_bytes, s);
            c.scratch(s).staging.copy(small + o_c, coeffs, (size_t)count * sizeof(Fr), s);
        }
        linear_combination_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const Fr* const*)small, (const Fr*)(small + o_c),
                                                                             count, n, (Fr*)d_out);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(s));   // the host arrays may go away after return
    });
}

}  // extern "C"
