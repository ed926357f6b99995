// `permute_expression_pair` of the lookup argument (SURVEY.md section 8 f, rank 1) on the device.
//
// Upstream ([DEP] halo2_proofs 0.2.0 @ v2023_01_20 src/plonk/lookup/prover.rs, reference
// Cargo.lock:469-471): over the usable rows m = n - (blinding_factors + 1),
//   permuted_input  = the input expression sorted by `Fr`'s `Ord` (canonical integer order);
//   permuted_table  = at every row where a value first appears in permuted_input, that value
//                     (one instance of it is taken out of the table's multiset; a missing value
//                     is Error::ConstraintSystemFailure); the remaining table values, ascending,
//                     fill the rows of repeated input values from the LAST such row backwards
//                     (`repeated_input_rows.pop()`);
//   the last blinding_factors + 1 rows of both are the caller's random scalars.
// The result is unique, so parity is bit-exact.
//
// Device plan, `count` lookups in one pipeline (blockIdx.y = lookup): convert to canonical
// integers, bitonic sort (tiles of 1024 keys in shared memory, wider strides in global memory),
// then flags + block scans + binary searches for the multiset difference.  This is bookkeeping
// around 254-bit compares (integer ALU work), not multiplier-bound like the rest of the path.
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ntt.cuh"

#include <vector>

namespace zk {

constexpr int SORT_TILE = 1024;       // keys per shared-memory tile
constexpr int SORT_THREADS = 512;

struct Key { uint32_t l[8]; };

__device__ __forceinline__ bool key_less(const Key& a, const Key& b) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a.l[i] != b.l[i]) return a.l[i] < b.l[i];
    }
    return false;
}
__device__ __forceinline__ bool key_eq(const Key& a, const Key& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}
__device__ __forceinline__ Key ld_key(const Key* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    Key r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_key(Key* p, const Key& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// keys[c][i] = canonical(src[c][i]) for i < m, +infinity (all ones) for m <= i < padded
static __global__ void sort_load_kernel(const Fr* __restrict__ src, size_t stride, uint32_t m, uint32_t padded,
                                        Key* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= padded) return;
    Key k;
    if (i < m) {
        const Fr v = ldg_fr(src + (size_t)c * stride + i).from_mont();
#pragma unroll
        for (int t = 0; t < 8; ++t) k.l[t] = v.l[t];
    } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) k.l[t] = 0xffffffffu;
    }
    st_key(keys + (size_t)c * padded + i, k);
}

// Bitonic network on one tile held in shared memory.  `k_first`..`k_last` are the merge sizes
// run here (all with stride j < SORT_TILE); `full` = start from unsorted data (k from 2).
static __global__ void __launch_bounds__(SORT_THREADS) sort_tile_kernel(Key* keys, uint32_t padded, uint32_t k_merge) {
    extern __shared__ uint4 sort_smem[];
    Key* sh = reinterpret_cast<Key*>(sort_smem);
    const uint32_t c = blockIdx.y, base = blockIdx.x * SORT_TILE;
    Key* col = keys + (size_t)c * padded;
    for (uint32_t t = threadIdx.x; t < SORT_TILE; t += SORT_THREADS)
        if (base + t < padded) sh[t] = ld_key(col + base + t);
    __syncthreads();
    const uint32_t tile = min((uint32_t)SORT_TILE, padded);
    // k_merge == 0: full sort of the tile; otherwise only the j < SORT_TILE stages of merge size k_merge
    for (uint32_t k = (k_merge ? k_merge : 2u); k <= (k_merge ? k_merge : tile); k <<= 1) {
        for (uint32_t j = min(k >> 1, tile >> 1); j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < tile / 2; t += SORT_THREADS) {
                const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const bool up = ((base + lo) & k) == 0;
                const Key a = sh[lo], b = sh[hi];
                if (key_less(b, a) == up) { sh[lo] = b; sh[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t t = threadIdx.x; t < SORT_TILE; t += SORT_THREADS)
        if (base + t < padded) st_key(col + base + t, sh[t]);
}

// One compare-exchange stage with stride j >= SORT_TILE of merge size k
static __global__ void sort_global_stage_kernel(Key* keys, uint32_t padded, uint32_t k, uint32_t j) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (t >= padded / 2) return;
    Key* col = keys + (size_t)c * padded;
    const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
    const bool up = (lo & k) == 0;
    const Key a = ld_key(col + lo), b = ld_key(col + hi);
    if (key_less(b, a) == up) { st_key(col + lo, b); st_key(col + hi, a); }
}

static void bitonic_sort(Key* keys, uint32_t padded, uint32_t count, cudaStream_t s) {
    const int smem = SORT_TILE * sizeof(Key);
    const uint32_t tiles = (padded + SORT_TILE - 1) / SORT_TILE;
    sort_tile_kernel<<<dim3(tiles, count), SORT_THREADS, smem, s>>>(keys, padded, 0);
    ZK_LAUNCH_CHECK();
    for (uint32_t k = 2 * SORT_TILE; k <= padded; k <<= 1) {
        for (uint32_t j = k >> 1; j >= (uint32_t)SORT_TILE; j >>= 1) {
            sort_global_stage_kernel<<<dim3((padded / 2 + 255) / 256, count), 256, 0, s>>>(keys, padded, k, j);
            ZK_LAUNCH_CHECK();
        }
        sort_tile_kernel<<<dim3(tiles, count), SORT_THREADS, smem, s>>>(keys, padded, k);
        ZK_LAUNCH_CHECK();
    }
}

// ---- multiset bookkeeping: one block per lookup, running offsets across 1024-element strips
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t sv = lane < nw ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, sv, o);
            if (lane >= (uint32_t)o) sv += y;
        }
        warp_sums[lane] = sv;
    }
    __syncthreads();
    const uint32_t basev = wid ? warp_sums[wid - 1] : 0u;
    total = warp_sums[nw - 1];
    __syncthreads();
    return basev + x - v;
}

__device__ __forceinline__ bool sorted_contains(const Key* a, uint32_t m, const Key& v) {
    uint32_t lo = 0, hi = m;              // first index with a[idx] >= v
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key_less(ld_key(a + mid), v)) lo = mid + 1; else hi = mid;
    }
    return lo < m && key_eq(ld_key(a + lo), v);
}

// A = sorted input, T = sorted table (canonical keys, `padded` per lookup, m usable).
// leftover[c][..] = T without one instance of every distinct value of A (ascending);
// rep_rank[c][i] = number of repeated rows before row i (valid where A[i] == A[i-1]);
// counts[c] = {#distinct values of A, #table entries removed, #repeated rows}.
static __global__ void __launch_bounds__(1024) permute_plan_kernel(const Key* __restrict__ A, const Key* __restrict__ T,
                                                                   uint32_t padded, uint32_t m, Key* __restrict__ leftover,
                                                                   uint32_t* __restrict__ rep_rank,
                                                                   uint32_t* __restrict__ counts) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t c = blockIdx.x;
    const Key* a = A + (size_t)c * padded;
    const Key* t = T + (size_t)c * padded;
    Key* lo_out = leftover + (size_t)c * padded;
    uint32_t* rr = rep_rank + (size_t)c * padded;
    uint32_t n_rep = 0, n_left = 0, n_removed = 0;
    for (uint32_t base = 0; base < m; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool repeated = false, keep = false, removed = false;
        Key tv;
        if (i < m) {
            const Key av = ld_key(a + i);
            repeated = i > 0 && key_eq(av, ld_key(a + i - 1));
            tv = ld_key(t + i);
            const bool first_in_table = i == 0 || !key_eq(tv, ld_key(t + i - 1));
            removed = first_in_table && sorted_contains(a, m, tv);
            keep = !removed;
        }
        uint32_t tot;
        const uint32_t r = block_scan_excl(repeated ? 1u : 0u, warp_sums, tot);
        if (i < m) rr[i] = n_rep + r;
        n_rep += tot;
        const uint32_t p = block_scan_excl(keep ? 1u : 0u, warp_sums, tot);
        if (keep) st_key(lo_out + n_left + p, tv);
        n_left += tot;
        (void)block_scan_excl(removed ? 1u : 0u, warp_sums, tot);
        n_removed += tot;
    }
    if (threadIdx.x == 0) {
        counts[3 * c] = m - n_rep;
        counts[3 * c + 1] = n_removed;
        counts[3 * c + 2] = n_rep;
    }
}

// Outputs in Montgomery form; rows >= m from the caller's blinding scalars (or zero)
static __global__ void permute_emit_kernel(const Key* __restrict__ A, const Key* __restrict__ leftover,
                                           const uint32_t* __restrict__ rep_rank, const uint32_t* __restrict__ counts,
                                           uint32_t padded, uint32_t m, uint32_t n, const Fr* __restrict__ blinds,
                                           Fr* __restrict__ out_input, Fr* __restrict__ out_table, size_t out_stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= n) return;
    Fr* oi = out_input + (size_t)c * out_stride + i;
    Fr* ot = out_table + (size_t)c * out_stride + i;
    if (i >= m) {
        const uint32_t nbl = n - m;
        Fr bi = Fr::zero(), bt = Fr::zero();
        if (blinds) {
            bi = ldg_fr(blinds + ((size_t)c * 2) * nbl + (i - m)).canon();
            bt = ldg_fr(blinds + ((size_t)c * 2 + 1) * nbl + (i - m)).canon();
        }
        st_fr(oi, bi);
        st_fr(ot, bt);
        return;
    }
    const Key* a = A + (size_t)c * padded;
    const Key av = ld_key(a + i);
    Fr x;
#pragma unroll
    for (int t = 0; t < 8; ++t) x.l[t] = av.l[t];
    st_fr(oi, x.to_mont().canon());
    const bool repeated = i > 0 && key_eq(av, ld_key(a + i - 1));
    Key tv = av;
    if (repeated) {
        const uint32_t n_rep = counts[3 * c + 2];
        tv = ld_key(leftover + (size_t)c * padded + (n_rep - 1 - rep_rank[(size_t)c * padded + i]));
    }
    Fr y;
#pragma unroll
    for (int t = 0; t < 8; ++t) y.l[t] = tv.l[t];
    st_fr(ot, y.to_mont().canon());
}

}  // namespace zk

using namespace zk;

extern "C" {

int b200zk_permute_expression_pair_dev(const void* d_input, const void* d_table, size_t stride, uint32_t count, uint32_t k,
                                       uint32_t blinding_factors, const uint64_t* blinds_or_null, void* d_permuted_input,
                                       void* d_permuted_table, size_t out_stride, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_input && d_table && d_permuted_input && d_permuted_table, "null argument");
        ZK_REQUIRE(count >= 1 && count <= 65535 && k >= 1 && k <= 26, "bad lookup shape");
        const uint32_t n = 1u << k;
        ZK_REQUIRE(blinding_factors + 1 < n, "too many blinding factors");
        ZK_REQUIRE(count == 1 || (stride >= n && out_stride >= n), "batch stride smaller than the column");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        const uint32_t m = n - (blinding_factors + 1);
        uint32_t padded = 1;
        while (padded < m) padded <<= 1;
        const size_t col_bytes = (size_t)padded * sizeof(Key);
        const size_t nbl = n - m;
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
        const size_t o_a = carve(col_bytes * count), o_t = carve(col_bytes * count), o_l = carve(col_bytes * count);
        const size_t o_rr = carve((size_t)padded * 4 * count), o_cnt = carve((size_t)count * 12);
        const size_t o_bl = carve(std::max<size_t>(1, (size_t)count * 2 * nbl) * sizeof(Fr));
        char* w = (char*)c.scratch(s).poly_cols.get(off);
        Key* A = (Key*)(w + o_a);
        Key* T = (Key*)(w + o_t);
        Key* L = (Key*)(w + o_l);
        uint32_t* rr = (uint32_t*)(w + o_rr);
        uint32_t* cnt = (uint32_t*)(w + o_cnt);
        Fr* bl = nullptr;
        if (blinds_or_null) {
            bl = (Fr*)(w + o_bl);
            c.scratch(s).staging.copy(bl, blinds_or_null, (size_t)count * 2 * nbl * sizeof(Fr), s);
        }
        static int configured_device = -1;      // per-device attribute (see ntt.cu launch_pass)
        if (configured_device != c.device) {
            ZK_CUDA(cudaFuncSetAttribute(sort_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(SORT_TILE * sizeof(Key))));
            configured_device = c.device;
        }
        const dim3 lgrid((padded + 255) / 256, count);
        sort_load_kernel<<<lgrid, 256, 0, s>>>((const Fr*)d_input, stride, m, padded, A);
        ZK_LAUNCH_CHECK();
        sort_load_kernel<<<lgrid, 256, 0, s>>>((const Fr*)d_table, stride, m, padded, T);
        ZK_LAUNCH_CHECK();
        bitonic_sort(A, padded, count, s);
        bitonic_sort(T, padded, count, s);
        permute_plan_kernel<<<count, 1024, 0, s>>>(A, T, padded, m, L, rr, cnt);
        ZK_LAUNCH_CHECK();
        std::vector<uint32_t> h(3 * (size_t)count);
        ZK_CUDA(cudaMemcpyAsync(h.data(), cnt, h.size() * 4, cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
        for (uint32_t j = 0; j < count; ++j)
            if (h[3 * j] != h[3 * j + 1])
                throw Error{"b200zk: lookup " + std::to_string(j) +
                            ": an input value is missing from the table (ConstraintSystemFailure)"};
        permute_emit_kernel<<<dim3((n + 255) / 256, count), 256, 0, s>>>(A, L, rr, cnt, padded, m, n, bl,
                                                                         (Fr*)d_permuted_input, (Fr*)d_permuted_table,
                                                                         out_stride);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
