// BN254 G1 multi-scalar multiplication (Pippenger) for sm_100a + its C ABI.
//
// Device replacement for halo2_proofs 0.2.0 `arithmetic::best_multiexp` and the two
// callers `ParamsKZG::commit` / `commit_lagrange` ([DEP] halo2_proofs/src/arithmetic.rs,
// halo2_proofs/src/poly/kzg/commitment.rs @ v2023_01_20, reference Cargo.lock:469-471).
// The result is the same group element sum_i coeffs[i] * bases[i]; upstream's unsigned
// ceil(ln n)-bit windows per rayon chunk are a CPU scheduling choice, not part of the
// result, and are not reproduced (oracle/ restates them for the CPU baseline).
//
// Pipeline (all on the device, one stream):
//   1. digits     scalar -> canonical (one Montgomery reduction) -> signed c-bit digits;
//                 histogram of (window, |digit|) keys.  Zero digits are dropped, so
//                 sparse / small witnesses cost proportionally less.
//   2. scan       exclusive prefix sum of the histogram -> bucket start offsets.
//   3. scatter    counting sort of (key, point index | sign) pairs by key.
//   4. accumulate every thread adds a fixed-length chunk of the sorted pair list into
//                 an XYZZ accumulator (mixed addition, 8M+2S), writing a bucket when
//                 its run ends inside the chunk and handing the open last run to the
//                 next level.  Work per thread is constant whatever the bucket sizes
//                 are, so skewed scalar distributions (all-equal, 0/1 witnesses) do
//                 not serialise.  Levels >= 1 repeat this on (key, XYZZ) partials.
//   5. reduce     per window sum_b b * B_b by segmented running sums, segment partials
//                 combined by the same keyed reduction (key = window).
//   6. fold       Horner over the windows with c doublings per step; Jacobian out.
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ec.cuh"
#include "ntt.cuh"  // ld_fr / st_fr helpers

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

namespace zk {

// Registered bases (ParamsKZG::g / g_lagrange): the points, and optionally the table
// T[w * n + i] = 2^(c*w) * P_i in affine form, which lets every window share one bucket set
// (no per-window reduction, no doubling chain at the end).
struct BaseTable {
    G1Affine* d = nullptr;
    size_t n = 0;
    G1Affine* table = nullptr;
    uint32_t c = 0, nwin = 0;
};

static void msm_release_pipeline();

static void msm_release_stage_events();
void msm_release_bases(Context& c) {
    msm_release_pipeline();
    msm_release_stage_events();
    for (auto& kv : c.bases) {
        cudaFree(kv.second->d);
        if (kv.second->table) cudaFree(kv.second->table);
        delete kv.second;
    }
    c.bases.clear();
}

// ------------------------------------------------------------------------- helpers
__device__ __forceinline__ Fq ldg_fq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
// 64-byte points gathered at random: ask L2 to fetch exactly the two sectors of the point
// (the default promotion pulls the whole 128-byte line, i.e. the neighbouring point as well).
__device__ __forceinline__ Fq ldg_fq_64B(const Fq* p) {
    uint4 a, b;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(reinterpret_cast<const uint4*>(p) + 1));
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fq ld_fq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fq(Fq* p, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ G1Xyzz ld_xyzz(const G1Xyzz* p) {
    G1Xyzz r;
    r.x = ld_fq(&p->x); r.y = ld_fq(&p->y); r.zz = ld_fq(&p->zz); r.zzz = ld_fq(&p->zzz);
    return r;
}
__device__ __forceinline__ void st_xyzz(G1Xyzz* p, const G1Xyzz& v) {
    st_fq(&p->x, v.x); st_fq(&p->y, v.y); st_fq(&p->zz, v.zz); st_fq(&p->zzz, v.zzz);
}

// Signed c-bit digits of a canonical 254-bit scalar; calls f(window, digit) for every
// non-zero digit, digit in [-2^(c-1), 2^(c-1)].
template <class F>
__device__ __forceinline__ void for_each_digit(const uint32_t* s, uint32_t c, uint32_t nwin, F&& f) {
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1u;
    uint32_t carry = 0;
    for (uint32_t w = 0; w < nwin; ++w) {
        const uint32_t o = w * c;
        const uint32_t limb = o >> 5, sh = o & 31u;
        uint32_t raw = 0;
        if (limb < 8) {
            raw = s[limb] >> sh;
            if (sh + c > 32 && limb + 1 < 8) raw |= s[limb + 1] << (32 - sh);
        }
        uint32_t d = (raw & mask) + carry;
        int32_t sd;
        if (d > half) { sd = (int32_t)d - (int32_t)(1u << c); carry = 1; }
        else { sd = (int32_t)d; carry = 0; }
        if (sd != 0) f(w, sd);
    }
}

// ------------------------------------------------------------ 1. digits + histogram
// Also stores the signed digits window-major (digits[(col * nwin + w) * n + i], 0 = dropped)
// so that the scatter can run one window at a time: all blocks in flight then touch one
// window's counters (2^(c-1) * 4 B) and one window's slice of the sorted list (<= n * 4 B),
// which stay in the 126 MB L2 instead of spraying 4-byte stores over every window at once.
// blockIdx.y = column of a batch.  key = (col * key_windows + (key_windows > 1 ? w : 0)) * nb
// + |digit| - 1: with a precomputed table all windows share one bucket set (key_windows = 1).
__global__ void msm_hist_kernel(const Fr* __restrict__ scalars, size_t scalar_stride, size_t n, uint32_t c,
                                uint32_t nwin, uint32_t key_windows, uint32_t* __restrict__ hist,
                                int32_t* __restrict__ digits) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;                     // no early return: the warp votes below need every lane
    const uint32_t col = blockIdx.y;
    Fr s = Fr::zero();
    if (valid) s = ldg_fr(scalars + (size_t)col * scalar_stride + i).from_mont();
    const uint32_t nb = 1u << (c - 1);
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t carry = 0;
    for (uint32_t w = 0; w < nwin; ++w) {
        const uint32_t o = w * c;
        const uint32_t limb = o >> 5, sh = o & 31u;
        uint32_t raw = 0;
        if (limb < 8) {
            raw = s.l[limb] >> sh;
            if (sh + c > 32 && limb + 1 < 8) raw |= s.l[limb + 1] << (32 - sh);
        }
        const uint32_t d = (raw & mask) + carry;
        int32_t sd;
        if (d > half) { sd = (int32_t)d - (int32_t)(1u << c); carry = 1; }
        else { sd = (int32_t)d; carry = 0; }
        if (valid) digits[((size_t)col * nwin + w) * n + i] = sd;
        // One atomic per distinct key in the warp: witness columns are full of repeated values
        // (all-equal scalars put every point of a window in one bucket), and same-address
        // atomics serialise.
        const uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd);
        const uint32_t key = (valid && sd != 0) ? mag - 1 : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (key != 0xffffffffu && lane == (uint32_t)(__ffs(peers) - 1)) {
            const size_t group = (size_t)col * key_windows + (key_windows > 1 ? w : 0);
            atomicAdd(hist + group * nb + key, (uint32_t)__popc(peers));
        }
    }
}

// ------------------------------------------------------------------- 2. prefix scan
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = (lane < blockDim.x / 32) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += y;
        }
        warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const uint32_t base = wid ? warp_sums[wid - 1] : 0u;
    total = warp_sums[blockDim.x / 32 - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  uint32_t* __restrict__ bsum, size_t n) {
    __shared__ uint32_t warp_sums[32];
    const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        t += v[i];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(t, warp_sums, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint32_t* bsum, uint32_t nblocks, uint32_t* total_out) {
    __shared__ uint32_t warp_sums[32];
    uint32_t running = 0;
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < nblocks) ? bsum[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_sums, total);
        if (i < nblocks) bsum[i] = running + ex;
        running += total;
    }
    if (threadIdx.x == 0) *total_out = running;
}

// What the pair count decides, computed on the device so that the host never has to wait for it:
// the accumulation's chunk length and thread count and the entry count of every keyed-reduction
// level.  The host launches grids for the worst case; threads beyond the planned counts exit.
constexpr int MSM_MAX_LEVELS = 12;
struct MsmRun {
    uint32_t npairs, L, nthreads;
    uint32_t any_long;   // some bucket has more open partial sums than the direct finish takes: run the keyed levels
    uint32_t level_count[MSM_MAX_LEVELS];   // entries entering combine level l (bucket accumulation)
    uint32_t level_count2[MSM_MAX_LEVELS];  // the same for the keyed reduction of the bucket-segment partials
};

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t* __restrict__ out, uint32_t* __restrict__ copy,
                                                                const uint32_t* __restrict__ bsum, size_t n,
                                                                const uint32_t* __restrict__ total, uint32_t chunk_unit,
                                                                uint32_t max_chunk, MsmRun* __restrict__ run) {
    const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    const uint32_t add = bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) {
            const uint32_t v = out[base + i] + add;
            out[base + i] = v;
            copy[base + i] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t np = *total;
        out[n] = np;  // start[K] = number of pairs
        // chunk length: keep >= chunk_unit threads queued, between 4 and max_chunk pairs each
        const uint32_t L = min(max_chunk, max(4u, np / chunk_unit));
        run->npairs = np;
        run->L = L;
        run->nthreads = (np + L - 1) / L;
        run->any_long = 0;
        run->level_count[0] = run->nthreads;
    }
}

// --------------------------------------------------------------------- 3. scatter
// grid = (ceil(n / 256), columns * windows * nsub): blocks of one (window, bucket sub-range)
// are scheduled together; the digits are re-read nsub times, streaming.  The 4-byte stores
// land at random places of the sorted list (ncu at k = 24: 9.8 GB of DRAM traffic for 1.6 GB
// of algorithmic bytes: every store is a sector read-modify-write); restricting a pass to a
// bucket sub-range narrows the set of lines being written.
// The stored value is the index of the point to add: i, or w * table_stride + i into the
// precomputed table, with the sign of the digit in bit 31.
__global__ void msm_scatter_kernel(const int32_t* __restrict__ digits, size_t n, uint32_t c, uint32_t nwin,
                                   uint32_t key_windows, uint32_t table_stride, uint32_t sub_bits,
                                   uint32_t index_offset, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cw = blockIdx.y >> sub_bits;             // col * nwin + w
    const uint32_t sub = blockIdx.y & ((1u << sub_bits) - 1u);
    const uint32_t lane = threadIdx.x & 31u;
    const int32_t d = (i < n) ? __ldcs(digits + (size_t)cw * n + i) : 0;
    const uint32_t mag = (uint32_t)(d < 0 ? -d : d);
    const bool mine = d != 0 && ((mag - 1) >> (c - 1 - sub_bits)) == sub;
    // warp-aggregated cursor bump: one atomic per distinct bucket in the warp
    const uint32_t key = mine ? mag - 1 : 0xffffffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
    const uint32_t col = cw / nwin, w = cw - col * nwin;
    const uint32_t nb = 1u << (c - 1);
    const size_t group = (size_t)col * key_windows + (key_windows > 1 ? w : 0);
    uint32_t base = 0;
    if (mine && lane == leader) base = atomicAdd(cursor + group * nb + key, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!mine) return;
    const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    sorted[pos] = ((uint32_t)i + index_offset + w * table_stride) | (d < 0 ? 0x80000000u : 0u);
}

// ------------------------------------------------ 1'-3'. two-level partition sort (large MSMs)
// The counting sort above bumps one L2 counter per pair twice (201 M `RED` + 201 M `ATOM` at 2^24: 1.6 + 5.5 ms,
// the second with a 4-byte store at a random place of a 805 MB list).  For large inputs the pairs are sorted in two
// levels instead, with every per-pair counter in shared memory:
//   A. keys are split into <= 4096 partitions by their high bits; a block ranks a tile of 8192 digits by partition in
//      shared memory, reserves room for each partition with one global atomic per (tile, partition) and writes
//      8-byte records (point | sign, low key bits) into the partition's region;
//   B. inside a partition (<= 2048 buckets) a block ranks a tile of records by bucket in shared memory — first only
//      counting, for the exact bucket offsets the accumulation needs, then again to place the 4-byte entries, one
//      global atomic per (tile, bucket).
// Global atomics drop from 2 per pair to ~0.4, the random 4-byte stores become runs, and the bucket offsets
// (`start`), the sorted list and the device plan come out exactly as from the counting sort.
// tiles of 32 entries per thread.  Measured at 2^24 (scratch A/B builds): pass A is fastest with 8192-entry tiles
// (256 threads, two blocks per SM), pass B with 16384-entry tiles (512 threads: runs of 8 entries per bucket instead
// of 4 in the ordered write-out): 1.49 -> 1.19 ms for B2, while the same tile made A 0.2 ms slower.
constexpr uint32_t PSORT_THREADS_A = 256, PSORT_TILE_A = 32 * PSORT_THREADS_A;
constexpr uint32_t PSORT_THREADS_B = 512, PSORT_TILE_B = 32 * PSORT_THREADS_B;
constexpr uint32_t PSORT_MAX_PARTS = 4096;
constexpr uint32_t PSORT_MAX_SHIFT = 12;     // <= 4096 buckets per partition

struct PsortArgs {
    uint32_t c, nwin, key_windows, shift, nparts;
    uint32_t table_stride, index_offset;
};

__device__ __forceinline__ bool psort_decode(const PsortArgs& A, int32_t d, uint32_t cw, uint32_t& key) {
    if (d == 0) return false;
    const uint32_t mag = (uint32_t)(d < 0 ? -d : d);
    const uint32_t col = cw / A.nwin, w = cw - col * A.nwin;
    const uint32_t group = col * A.key_windows + (A.key_windows > 1 ? w : 0);
    key = (group << (A.c - 1)) + mag - 1;
    return true;
}

// digits (as msm_hist_kernel) + the number of pairs of every partition
__global__ void __launch_bounds__(256) msm_digits_count_kernel(const Fr* __restrict__ scalars, size_t scalar_stride, size_t n,
                                                               PsortArgs A, uint32_t* __restrict__ part_count,
                                                               int32_t* __restrict__ digits) {
    __shared__ uint32_t sh[PSORT_MAX_PARTS];
    for (uint32_t t = threadIdx.x; t < A.nparts; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const uint32_t col = blockIdx.y;
    const uint32_t c = A.c, nwin = A.nwin;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1u;
#pragma unroll 1
    for (uint32_t j = 0; j < 8; ++j) {
        const size_t i = ((size_t)blockIdx.x * 8 + j) * blockDim.x + threadIdx.x;
        if (i >= n) break;
        const Fr s = ldg_fr(scalars + (size_t)col * scalar_stride + i).from_mont();
        uint32_t carry = 0;
        for (uint32_t w = 0; w < nwin; ++w) {
            const uint32_t o = w * c;
            const uint32_t limb = o >> 5, shb = o & 31u;
            uint32_t raw = 0;
            if (limb < 8) {
                raw = s.l[limb] >> shb;
                if (shb + c > 32 && limb + 1 < 8) raw |= s.l[limb + 1] << (32 - shb);
            }
            const uint32_t d = (raw & mask) + carry;
            int32_t sd;
            if (d > half) { sd = (int32_t)d - (int32_t)(1u << c); carry = 1; }
            else { sd = (int32_t)d; carry = 0; }
            digits[((size_t)col * nwin + w) * n + i] = sd;
            uint32_t key;
            if (psort_decode(A, sd, col * nwin + w, key)) atomicAdd(&sh[key >> A.shift], 1u);
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < A.nparts; t += blockDim.x)
        if (sh[t]) atomicAdd(part_count + t, sh[t]);
}

// exclusive scans of the partition sizes (-> part_start, a copy as cursors) and of their tile counts (-> tile_prefix)
__global__ void __launch_bounds__(1024) msm_partition_scan_kernel(const uint32_t* __restrict__ part_count, uint32_t nparts,
                                                                  uint32_t* __restrict__ part_start, uint32_t* __restrict__ part_cursor,
                                                                  uint32_t* __restrict__ tile_prefix) {
    __shared__ uint32_t sa[PSORT_MAX_PARTS + 1], sb[PSORT_MAX_PARTS + 1];
    for (uint32_t t = threadIdx.x; t < nparts; t += blockDim.x) {
        const uint32_t v = part_count[t];
        sa[t] = v;
        sb[t] = (v + PSORT_TILE_B - 1) / PSORT_TILE_B;
    }
    __syncthreads();
    if (threadIdx.x < 2) {                   // two serial scans of <= 4096 entries, side by side
        uint32_t* a = threadIdx.x == 0 ? sa : sb;
        uint32_t run = 0;
        for (uint32_t t = 0; t < nparts; ++t) { const uint32_t v = a[t]; a[t] = run; run += v; }
        a[nparts] = run;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t <= nparts; t += blockDim.x) {
        part_start[t] = sa[t];
        if (t < nparts) part_cursor[t] = sa[t];
        tile_prefix[t] = sb[t];
    }
}

// Block-wide exclusive scan of cnt[0..nbins) in place (nbins <= 4096, 256 threads: up to 16 consecutive bins each),
// then one global reservation per non-empty bin: delta[bin] = (position reserved in `cursor`) - (local position),
// so that the element at local sorted position q of bin `bin` goes to global position delta[bin] + q.
// `rot` rotates which bins a thread reserves, so that concurrent blocks do not hit the same counters in step.
// Returns the number of elements of the tile.  tot: 257 words of shared memory.
template <uint32_t THREADS>
__device__ __forceinline__ uint32_t psort_scan_reserve(uint32_t* cnt, uint32_t* delta, uint32_t nbins, uint32_t* cursor,
                                                       uint32_t rot, uint32_t* tot) {
    constexpr uint32_t PSORT_THREADS = THREADS;
    const uint32_t per = (nbins + PSORT_THREADS - 1) / PSORT_THREADS;      // <= 16 (4096 bins / 256 threads)
    const uint32_t t = (threadIdx.x + rot) % PSORT_THREADS;                 // logical slot of this thread
    const uint32_t b0 = t * per;
    uint32_t local[16];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) {
        local[j] = (j < per && b0 + j < nbins) ? cnt[b0 + j] : 0u;
        sum += local[j];
    }
    tot[t] = sum;
    __syncthreads();
    constexpr int SLOTS = PSORT_THREADS / 32;
    if (threadIdx.x < 32) {                // one warp scans the slot sums, SLOTS each
        uint32_t v[SLOTS], s8 = 0;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) { v[j] = tot[threadIdx.x * SLOTS + j]; s8 += v[j]; }
        uint32_t x = s8;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (threadIdx.x >= (uint32_t)o) x += y;
        }
        if (threadIdx.x == 31) tot[PSORT_THREADS] = x;
        uint32_t run = x - s8;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) { tot[threadIdx.x * SLOTS + j] = run; run += v[j]; }
    }
    __syncthreads();
    uint32_t run = tot[t];
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) {
        if (j < per && b0 + j < nbins) {
            cnt[b0 + j] = run;                                             // local exclusive offset
            delta[b0 + j] = local[j] ? atomicAdd(cursor + b0 + j, local[j]) - run : 0u;
            run += local[j];
        }
    }
    __syncthreads();
    return tot[PSORT_THREADS];
}

// A: digits -> records grouped by partition.  The tile is ordered by partition in shared memory first, so that
// consecutive threads store consecutive records of a partition's run: a scattered 8-byte store per lane is a memory
// request each, and 201 M of them were most of the first version's 4.1 ms.
__global__ void __launch_bounds__(PSORT_THREADS_A, 2) msm_partition_scatter_kernel(const int32_t* __restrict__ digits, size_t n,
                                                                                 uint64_t total, PsortArgs A,
                                                                                 uint32_t* __restrict__ part_cursor,
                                                                                 uint2* __restrict__ rec) {
    extern __shared__ uint4 psort_smem[];
    uint2* sh_rec = reinterpret_cast<uint2*>(psort_smem);                  // PSORT_TILE_A records
    uint32_t* sh_cnt = reinterpret_cast<uint32_t*>(sh_rec + PSORT_TILE_A);   // PSORT_MAX_PARTS
    uint32_t* sh_delta = sh_cnt + PSORT_MAX_PARTS;                         // PSORT_MAX_PARTS
    uint32_t* sh_tot = sh_delta + PSORT_MAX_PARTS;                         // PSORT_THREADS_A + 1
    for (uint32_t t = threadIdx.x; t < A.nparts; t += blockDim.x) sh_cnt[t] = 0;
    __syncthreads();
    constexpr int PER = PSORT_TILE_A / PSORT_THREADS_A;
    uint32_t pr[PER], val[PER], kl[PER];
    const uint64_t base = (uint64_t)blockIdx.x * PSORT_TILE_A;
    // (column, window) row and position of the tile's first entry; entries advance without divisions
    const uint32_t cw0 = (uint32_t)(base / n);
    const uint64_t i0 = base - (uint64_t)cw0 * n;
    const uint32_t lowmask = (1u << A.shift) - 1u;
    // all loads first: a shared-memory atomic between two loads keeps the compiler from overlapping them, and 32
    // dependent round trips to DRAM per thread were 60 % of this kernel's stall samples
    int32_t dv[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint64_t e = base + (uint64_t)j * PSORT_THREADS_A + threadIdx.x;
        dv[j] = e < total ? __ldcs(digits + e) : 0;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        uint64_t i = i0 + (uint64_t)j * PSORT_THREADS_A + threadIdx.x;
        uint32_t cw = cw0;
        while (i >= n) { i -= n; ++cw; }
        const int32_t d = dv[j];
        pr[j] = 0xffffffffu;
        uint32_t key;
        if (psort_decode(A, d, cw, key)) {
            const uint32_t p = key >> A.shift;
            const uint32_t w = cw % A.nwin;
            val[j] = ((uint32_t)i + A.index_offset + w * A.table_stride) | (d < 0 ? 0x80000000u : 0u);
            kl[j] = (key & lowmask) | (p << 16);
            pr[j] = (p << 16) | atomicAdd(&sh_cnt[p], 1u);
        }
    }
    __syncthreads();
    const uint32_t pop = psort_scan_reserve<PSORT_THREADS_A>(sh_cnt, sh_delta, A.nparts, part_cursor, blockIdx.x * 7u, sh_tot);
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (pr[j] != 0xffffffffu) sh_rec[sh_cnt[pr[j] >> 16] + (pr[j] & 0xffffu)] = make_uint2(val[j], kl[j]);
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < pop; q += PSORT_THREADS_A) {
        const uint2 r = sh_rec[q];
        rec[sh_delta[r.y >> 16] + q] = make_uint2(r.x, r.y & 0xffffu);
    }
}

// the partition and record range of a block of pass B
__device__ __forceinline__ bool psort_tile(const uint32_t* __restrict__ tile_prefix, const uint32_t* __restrict__ part_start,
                                           uint32_t nparts, uint32_t& p, uint32_t& lo, uint32_t& hi) {
    const uint32_t b = blockIdx.x;
    if (b >= __ldg(tile_prefix + nparts)) return false;
    uint32_t l = 0, h = nparts;                       // tile_prefix[l] <= b < tile_prefix[h]
    while (h - l > 1) {
        const uint32_t m = (l + h) >> 1;
        if (__ldg(tile_prefix + m) <= b) l = m; else h = m;
    }
    p = l;
    lo = __ldg(part_start + p) + (b - __ldg(tile_prefix + p)) * PSORT_TILE_B;
    hi = min(lo + PSORT_TILE_B, __ldg(part_start + p + 1));
    return true;
}

// B1: exact size of every bucket
__global__ void __launch_bounds__(PSORT_THREADS_B) msm_bucket_count_kernel(const uint2* __restrict__ rec, PsortArgs A,
                                                                         const uint32_t* __restrict__ tile_prefix,
                                                                         const uint32_t* __restrict__ part_start,
                                                                         uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[1u << PSORT_MAX_SHIFT];
    uint32_t p, lo, hi;
    if (!psort_tile(tile_prefix, part_start, A.nparts, p, lo, hi)) return;
    const uint32_t nbins = 1u << A.shift;
    for (uint32_t t = threadIdx.x; t < nbins; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    for (uint32_t e = lo + threadIdx.x; e < hi; e += blockDim.x) atomicAdd(&sh[__ldg(&rec[e].y)], 1u);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < nbins; t += blockDim.x)
        if (sh[t]) atomicAdd(hist + ((size_t)p << A.shift) + t, sh[t]);
}

// B2: records -> the sorted list of (point | sign) entries, ordered by bucket in shared memory first (see A)
__global__ void __launch_bounds__(PSORT_THREADS_B, 1) msm_bucket_scatter_kernel(const uint2* __restrict__ rec, PsortArgs A,
                                                                              const uint32_t* __restrict__ tile_prefix,
                                                                              const uint32_t* __restrict__ part_start,
                                                                              uint32_t* __restrict__ cursor,
                                                                              uint32_t* __restrict__ sorted) {
    extern __shared__ uint4 psort_smem[];
    uint2* sh_rec = reinterpret_cast<uint2*>(psort_smem);                  // PSORT_TILE_B x (value, bucket)
    uint32_t* sh_cnt = reinterpret_cast<uint32_t*>(sh_rec + PSORT_TILE_B);   // 2^PSORT_MAX_SHIFT
    uint32_t* sh_delta = sh_cnt + (1u << PSORT_MAX_SHIFT);
    uint32_t* sh_tot = sh_delta + (1u << PSORT_MAX_SHIFT);
    uint32_t p, lo, hi;
    if (!psort_tile(tile_prefix, part_start, A.nparts, p, lo, hi)) return;
    const uint32_t nbins = 1u << A.shift;
    for (uint32_t t = threadIdx.x; t < nbins; t += blockDim.x) sh_cnt[t] = 0;
    __syncthreads();
    constexpr int PER = PSORT_TILE_B / PSORT_THREADS_B;
    uint32_t val[PER], kr[PER];
    uint32_t kb[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {          // all loads first (see msm_partition_scatter_kernel)
        const uint32_t e = lo + (uint32_t)j * PSORT_THREADS_B + threadIdx.x;
        const uint2 r = e < hi ? __ldcs(rec + e) : make_uint2(0u, 0xffffffffu);
        val[j] = r.x;
        kb[j] = r.y;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        kr[j] = 0xffffffffu;
        if (kb[j] != 0xffffffffu) kr[j] = (kb[j] << 16) | atomicAdd(&sh_cnt[kb[j]], 1u);
    }
    __syncthreads();
    const uint32_t pop = psort_scan_reserve<PSORT_THREADS_B>(sh_cnt, sh_delta, nbins, cursor + ((size_t)p << A.shift), blockIdx.x * 7u, sh_tot);
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (kr[j] != 0xffffffffu) sh_rec[sh_cnt[kr[j] >> 16] + (kr[j] & 0xffffu)] = make_uint2(val[j], kr[j] >> 16);
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < pop; q += PSORT_THREADS_B) {
        const uint2 r = sh_rec[q];
        sorted[sh_delta[r.y] + q] = r.x;
    }
}

constexpr uint32_t MSM_PAD_KEY = 0xffffffffu;   // padding lane (real keys are < 2^31)

// ------------------------------------------------------- 4. bucket accumulation, level 0
// start[0..K] are the bucket offsets (start[K] = npairs).  Thread t owns pairs
// [t*L, (t+1)*L).  A run that ends inside the chunk is complete on its right side and is
// stored to buckets[key] (each key has exactly one such writer per level; buckets were
// cleared to the identity).  The last run of the chunk is always handed up as
// (carry_key[t], carry_pt[t]).
__device__ __forceinline__ uint32_t find_key(const uint32_t* __restrict__ start, uint32_t nkeys, uint32_t pos) {
    // largest key with start[key] <= pos  (start[key+1] > pos)
    uint32_t lo = 0, hi = nkeys;  // invariant: start[lo] <= pos < start[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(start + mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// (Tried in round 2: short chunks (L / 4) for the last machine-full of pairs, so that the kernel drains in a quarter
// of a chunk's time: accumulate 32.10 -> 31.77 ms at 2^24, but the 0.38 M extra open runs cost the keyed levels
// 0.73 -> 1.08 ms.  Not kept.)
// (Tried in round 2: the ten products of the mixed addition as calls of one __noinline__ multiplication / squaring
// body instead of ten inlined copies, to shrink the ~37 KB loop body: 2 950 instead of 4 300 SASS instructions, but
// 244 MOV + 226 SEL of argument marshalling and 2x the spills: 33.9 ms against 32.1 ms at 2^24.  Not kept.)
// ADD = false: buckets were cleared, a closed run is stored.  ADD = true (a later point range of the
// same MSM, b200zk_msm_g1_registered's upload pipeline): a run that closes inside the chunk
// continues from what the earlier ranges left in the bucket.
template <bool ADD>
__global__ void __launch_bounds__(128, 5) msm_accum_kernel(const G1Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                        const uint32_t* __restrict__ start, uint32_t nkeys,
                                                        const MsmRun* __restrict__ run, G1Xyzz* __restrict__ buckets,
                                                        uint32_t* __restrict__ carry_key, G1Xyzz* __restrict__ carry_pt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t npairs = __ldg(&run->npairs), L = __ldg(&run->L);
    if (t >= __ldg(&run->nthreads)) return;
    const uint32_t begin = t * L;
    const uint32_t end = min(begin + L, npairs);
    uint32_t key = find_key(start, nkeys, begin);
    uint32_t run_end = __ldg(start + key + 1);
    // ADD: the thread that will close a run (the only level-0 writer of that bucket) starts from
    // what the earlier point ranges left there, so the hot loop carries no extra addition
    G1Xyzz acc = (ADD && run_end <= end) ? ld_xyzz(buckets + key) : G1Xyzz::identity();
    // software prefetch of the next point
    uint32_t e = __ldg(sorted + begin);
    G1Affine nxt;
    nxt.x = ldg_fq_64B(&bases[e & 0x7fffffffu].x);
    nxt.y = ldg_fq_64B(&bases[e & 0x7fffffffu].y);
#pragma unroll 1
    for (uint32_t pos = begin; pos < end; ++pos) {
        G1Affine p = nxt;
        const bool negate = (e >> 31) != 0;
        if (pos + 1 < end) {
            e = __ldg(sorted + pos + 1);
            nxt.x = ldg_fq_64B(&bases[e & 0x7fffffffu].x);
            nxt.y = ldg_fq_64B(&bases[e & 0x7fffffffu].y);
        }
        if (pos >= run_end) {
            st_xyzz(buckets + key, acc);
            do {
                ++key;
                run_end = __ldg(start + key + 1);
            } while (pos >= run_end);
            acc = (ADD && run_end <= end) ? ld_xyzz(buckets + key) : G1Xyzz::identity();
        }
        if (negate) p.y = p.y.neg();
        acc.add_affine(p);
    }
    if (run_end == end) {
        // the chunk ends exactly where its last run ends: nothing continues into the next
        // chunk, so the run is complete on its right side like the interior ones
        st_xyzz(buckets + key, acc);
        acc = G1Xyzz::identity();   // hand up (key, identity): keys stay dense and sorted
    }
    carry_key[t] = key;
    st_xyzz(carry_pt + t, acc);
}

// ------------------------------------------------------ 4b. keyed reduction, level >= 1
// Entries (keys[i], pts[i]) with non-decreasing keys.  A run of equal keys is *closed* where
// it ends (the next entry has another key, or there is none); exactly one thread per level
// sees a given key close, and only that thread adds the run's sum into buckets[key].  Open
// last runs are handed up as (key, partial); closed ones hand up (key, identity) so the next
// level's keys stay dense and sorted.
//
// Sequential form (throughput regime): thread t owns entries [t*L, (t+1)*L).
__global__ void __launch_bounds__(128) msm_combine_kernel(const uint32_t* __restrict__ keys, const G1Xyzz* __restrict__ pts,
                                                          uint32_t* __restrict__ counts, uint32_t level, uint32_t L,
                                                          G1Xyzz* __restrict__ buckets,
                                                          uint32_t* __restrict__ carry_key, G1Xyzz* __restrict__ carry_pt,
                                                          const uint32_t* __restrict__ enabled) {
    if (enabled && !*enabled) return;                // the direct finish has already closed every run
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = counts[level];
    const uint32_t nthreads = (count + L - 1) / L;
    if (t == 0) counts[level + 1] = nthreads;
    if (t >= nthreads) return;
    const uint32_t begin = t * L;
    const uint32_t end = min(begin + L, count);
    uint32_t key = keys[begin];
    G1Xyzz acc = G1Xyzz::identity();
    auto flush = [&]() {
        if (!acc.is_identity()) {
            G1Xyzz b = ld_xyzz(buckets + key);
            b.add(acc);
            st_xyzz(buckets + key, b);
        }
    };
#pragma unroll 1
    for (uint32_t i = begin; i < end; ++i) {
        const uint32_t k = keys[i];
        if (k != key) {
            flush();
            acc = G1Xyzz::identity();
            key = k;
        }
        G1Xyzz p = ld_xyzz(pts + i);
        acc.add(p);
    }
    if (end == count || keys[end] != key) {
        flush();
        acc = G1Xyzz::identity();
    }
    carry_key[t] = key;
    st_xyzz(carry_pt + t, acc);
}

__device__ __forceinline__ G1Xyzz shfl_down_xyzz(const G1Xyzz& v, uint32_t d) {
    G1Xyzz r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], d);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], d);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], d);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], d);
    }
    return r;
}

// Warp form (latency regime): one entry per lane, segmented shuffle reduction by key in five
// steps; each warp hands up one entry, so a level shrinks the list 32x for the latency of
// five point additions.  When count <= 32 the single warp closes everything.
__global__ void __launch_bounds__(128) msm_combine_warp_kernel(const uint32_t* __restrict__ keys,
                                                               const G1Xyzz* __restrict__ pts, uint32_t* __restrict__ counts,
                                                               uint32_t level, G1Xyzz* __restrict__ buckets,
                                                               uint32_t* __restrict__ carry_key,
                                                               G1Xyzz* __restrict__ carry_pt,
                                                               const uint32_t* __restrict__ enabled) {
    if (enabled && !*enabled) return;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = counts[level];
    if (gid == 0) counts[level + 1] = (count + 31) / 32;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = gid >> 5;
    const uint32_t base = warp << 5;
    if (base >= count) return;                       // whole warp beyond the list
    const bool valid = gid < count;
    const uint32_t key = valid ? keys[gid] : MSM_PAD_KEY;
    G1Xyzz v = valid ? ld_xyzz(pts + gid) : G1Xyzz::identity();
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t ok = __shfl_down_sync(0xffffffffu, key, d);
        const G1Xyzz o = shfl_down_xyzz(v, d);
        if (lane + d < 32 && ok == key) v.add(o);
    }
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const uint32_t last_lane = min(31u, count - base - 1u);
    const uint32_t k_last = __shfl_sync(0xffffffffu, key, last_lane);
    if (!(head && valid)) return;
    const bool reaches_end = (key == k_last);
    const bool open = reaches_end && (base + 32 < count) && (keys[base + 32] == key);
    if (open) {
        carry_key[warp] = key;
        st_xyzz(carry_pt + warp, v);
        return;
    }
    if (!v.is_identity()) {
        G1Xyzz b = ld_xyzz(buckets + key);
        b.add(v);
        st_xyzz(buckets + key, b);
    }
    if (reaches_end) {
        carry_key[warp] = key;
        st_xyzz(carry_pt + warp, G1Xyzz::identity());
    }
}

// Quad form (latency regime, see ec.cuh): one entry per quad, eight per warp; three segmented shuffle
// steps of 4-deep cooperative additions close every run that ends inside the warp.  A level shrinks
// the list 8x for the latency of ~4 short additions; a few thousand entries are gone in four levels.
__global__ void __launch_bounds__(128) msm_combine_quad_kernel(const uint32_t* __restrict__ keys,
                                                               const G1Xyzz* __restrict__ pts, uint32_t* __restrict__ counts,
                                                               uint32_t level, G1Xyzz* __restrict__ buckets,
                                                               uint32_t* __restrict__ carry_key,
                                                               G1Xyzz* __restrict__ carry_pt,
                                                               const uint32_t* __restrict__ enabled) {
    if (enabled && !*enabled) return;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t role = threadIdx.x & 3u, qlane = (threadIdx.x & 31u) >> 2;
    const uint32_t count = counts[level];
    if (gid == 0) counts[level + 1] = (count + 7) / 8;
    const uint32_t warp = gid >> 5;
    const uint32_t base = warp << 3;
    if (base >= count) return;                       // whole warp beyond the list
    const uint32_t e = base + qlane;
    const bool valid = e < count;
    const uint32_t key = valid ? keys[e] : MSM_PAD_KEY;
    G1Xyzz v = valid ? ld_xyzz(pts + e) : G1Xyzz::identity();
#pragma unroll 1
    for (uint32_t d = 1; d < 8; d <<= 1) {
        const uint32_t ok = __shfl_down_sync(0xffffffffu, key, 4 * d);
        const G1Xyzz o = shfl_down_xyzz(v, 4 * d);
        const bool take = (qlane + d < 8) && ok == key;
        v = quad_add(v, xyzz_sel(take, o, G1Xyzz::identity()), role);
    }
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 4);
    const bool head = valid && ((qlane == 0) || (prev != key));
    const uint32_t last_q = min(7u, count - base - 1u);
    const uint32_t k_last = __shfl_sync(0xffffffffu, key, last_q * 4);
    const bool reaches_end = (key == k_last);
    const bool open = head && reaches_end && (base + 8 < count) && (keys[base + 8] == key);
    const bool flush = head && !open && !v.is_identity();
    if (__any_sync(0xffffffffu, flush)) {
        G1Xyzz b = flush ? ld_xyzz(buckets + key) : G1Xyzz::identity();
        b = quad_add(b, xyzz_sel(flush, v, G1Xyzz::identity()), role);
        if (flush && role == 0) st_xyzz(buckets + key, b);
    }
    if (head && reaches_end && role == 0) {
        carry_key[warp] = key;
        st_xyzz(carry_pt + warp, open ? v : G1Xyzz::identity());
    }
}

// Direct finish of the bucket accumulation (one quad per bucket).  The open partial sums the accumulation
// handed up for bucket b sit at consecutive thread indices that follow from the bucket offsets alone:
// thread t's last pair is at position min((t + 1) L, npairs) - 1, so bucket [s, e) owns t in [s / L, e / L)
// (through the last thread when e = npairs).  With uniform scalars that is a handful per bucket; the quad
// adds them up and folds them into the bucket, and no keyed-reduction level has to run.  A bucket with more
// than `rmax` partials (skewed scalars: every point of a window in one bucket) raises run->any_long and is left to
// the keyed levels, which then run over the whole list; the partials consumed here are cleared so that they
// count once.
__global__ void __launch_bounds__(128) msm_finish_quad_kernel(const uint32_t* __restrict__ start, uint32_t nkeys,
                                                              MsmRun* __restrict__ run, uint32_t rmax,
                                                              G1Xyzz* __restrict__ buckets, G1Xyzz* __restrict__ carry_pt) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t role = threadIdx.x & 3u;
    const uint32_t q = gid >> 2;
    const bool live = q < nkeys;
    const uint32_t b = live ? q : nkeys - 1;
    const uint32_t L = run->L, npairs = run->npairs, nthreads = run->nthreads;
    const uint32_t s = __ldg(start + b), e = __ldg(start + b + 1);
    const uint32_t t0 = s / L, t1 = (e == npairs) ? nthreads : e / L;
    uint32_t r = (live && e > s) ? t1 - t0 : 0u;
    if (r > rmax) {
        if (role == 0) run->any_long = 1u;
        r = 0;
    }
    const uint32_t steps = __reduce_max_sync(0xffffffffu, r);
    G1Xyzz acc = G1Xyzz::identity();
#pragma unroll 1
    for (uint32_t i = 0; i < steps; ++i) {
        const bool take = i < r;
        G1Xyzz x = take ? ld_xyzz(carry_pt + t0 + i) : G1Xyzz::identity();
        acc = quad_add(acc, x, role);
        if (take && role == 0) st_fq(&carry_pt[t0 + i].zz, Fq::zero());      // consumed
    }
    const bool need = r > 0 && !acc.is_identity();
    if (__any_sync(0xffffffffu, need)) {
        G1Xyzz bk = need ? ld_xyzz(buckets + b) : G1Xyzz::identity();
        bk = quad_add(bk, acc, role);
        if (need && role == 0) st_xyzz(buckets + b, bk);
    }
}

// --------------------------------------------------------------- 5. window reduction
// Thread (w, seg) walks `seglen` buckets of window w from the top: run += B_b,
// sum += run, then emits  sum + (lo-1) * run  with key w, where lo is the weight of the
// segment's lowest bucket.  The sum over a window's segments is sum_b b * B_b.
__global__ void __launch_bounds__(128, 3) msm_reduce_kernel(const G1Xyzz* __restrict__ buckets, uint32_t nb, uint32_t seglen,
                                                         uint32_t segs_per_win, uint32_t nwin,
                                                         uint32_t* __restrict__ out_key, G1Xyzz* __restrict__ out_pt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= segs_per_win * nwin) return;
    const uint32_t w = t / segs_per_win, seg = t % segs_per_win;
    const uint32_t lo_idx = seg * seglen;                 // bucket index (weight = index + 1)
    const uint32_t hi_idx = min(lo_idx + seglen, nb);
    const G1Xyzz* B = buckets + (size_t)w * nb;
    G1Xyzz run = G1Xyzz::identity(), sum = G1Xyzz::identity();
#pragma unroll 1
    for (uint32_t b = hi_idx; b > lo_idx; --b) {
        G1Xyzz x = ld_xyzz(B + (b - 1));
        run.add(x);
        sum.add(run);
    }
    // sum = sum_{b} (b - lo_idx) * B_b  with weights 1..seglen; add lo_idx * run
    uint32_t k = lo_idx;
    if (k != 0 && !run.is_identity()) {
        G1Xyzz acc = G1Xyzz::identity();
        for (int bit = 31 - __clz(k); bit >= 0; --bit) {
            acc = acc.dbl();
            if ((k >> bit) & 1u) acc.add(run);
        }
        sum.add(acc);
    }
    out_key[t] = w;
    st_xyzz(out_pt + t, sum);
}


// The same reduction with one *quad* per segment (ec.cuh): every running-sum step and every step of the
// lo * run doubling chain is 3-4 multiplications deep instead of 9-14, and four times as many warps
// share the work, which is what this latency-bound stage lacks.  All quads of a warp run the same
// number of steps; buckets beyond the end read as the identity.
__global__ void __launch_bounds__(128) msm_reduce_quad_kernel(const G1Xyzz* __restrict__ buckets, uint32_t nb, uint32_t seglen,
                                                              uint32_t segs_per_win, uint32_t nwin, uint32_t lo_bits,
                                                              uint32_t* __restrict__ counts,
                                                              uint32_t* __restrict__ out_key, G1Xyzz* __restrict__ out_pt) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t role = threadIdx.x & 3u;
    const uint32_t total = segs_per_win * nwin;
    if (gid == 0) counts[0] = total;
    const uint32_t q = gid >> 2;
    const bool live = q < total;
    const uint32_t t = live ? q : total - 1;
    const uint32_t w = t / segs_per_win, seg = t % segs_per_win;
    const uint32_t lo_idx = seg * seglen;                 // bucket index (weight = index + 1)
    const G1Xyzz* B = buckets + (size_t)w * nb;
    G1Xyzz run = G1Xyzz::identity(), sum = G1Xyzz::identity();
#pragma unroll 1
    for (uint32_t i = seglen; i > 0; --i) {
        const uint32_t idx = lo_idx + i - 1;
        G1Xyzz x = (idx < nb) ? ld_xyzz(B + idx) : G1Xyzz::identity();
        run = quad_add(run, x, role);
        sum = quad_add(sum, run, role);
    }
    // sum = sum_{b} (b - lo_idx) * B_b  with weights 1..seglen; add lo_idx * run
    G1Xyzz acc = G1Xyzz::identity();
#pragma unroll 1
    for (int bit = (int)lo_bits - 1; bit >= 0; --bit) {
        if (__any_sync(0xffffffffu, !acc.is_identity())) acc = quad_dbl(acc, role);
        const bool set = ((lo_idx >> bit) & 1u) != 0;
        if (__any_sync(0xffffffffu, set)) {
            const G1Xyzz tsum = quad_add(acc, run, role);
            acc = xyzz_sel(set, tsum, acc);
        }
    }
    sum = quad_add(sum, acc, role);
    if (live && role == 0) {
        out_key[q] = w;
        st_xyzz(out_pt + q, sum);
    }
}

// ----------------------------------------------------------------------- 6. fold
// One quad per column of the batch: Horner over that column's windows.  Without a window table this is a chain
// of (windows - 1) * c doublings (≈ 260 at 2^24, 0.55 ms on one lane): the 4-lane cooperative doubling is three
// products deep instead of nine.
__global__ void __launch_bounds__(32) msm_fold_kernel(const G1Xyzz* __restrict__ win, uint32_t nwin, uint32_t c, uint32_t count,
                                                      G1Jacobian* out) {
    const uint32_t role = threadIdx.x & 3u;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = q < count;
    const uint32_t col = live ? q : count - 1;
    G1Xyzz acc = G1Xyzz::identity();
    for (int w = (int)nwin - 1; w >= 0; --w) {
        if (__any_sync(0xffffffffu, !acc.is_identity()))
            for (uint32_t i = 0; i < c; ++i) acc = quad_dbl(acc, role);
        G1Xyzz x = ld_xyzz(win + (size_t)col * nwin + w);
        acc = quad_add(acc, x, role);
    }
    if (live && role == 0) out[col] = acc.to_jacobian();
}

// T[w * n + i] = 2^(c*w) * P_i, affine.  One thread per point walks the doubling chain and
// normalises each multiple with a field inversion (one-off cost at registration).
__global__ void __launch_bounds__(128) msm_precompute_kernel(const G1Affine* __restrict__ bases, size_t n, uint32_t c,
                                                             uint32_t nwin, G1Affine* __restrict__ table) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = ldg_fq(&bases[i].x);
    p.y = ldg_fq(&bases[i].y);
    st_fq(&table[i].x, p.x);
    st_fq(&table[i].y, p.y);
    for (uint32_t w = 1; w < nwin; ++w) {
        if (!p.is_identity()) {
            G1Xyzz a = G1Xyzz::double_affine(p);
            for (uint32_t k = 1; k < c; ++k) a = a.dbl();
            G1Jacobian j = a.to_jacobian_normalized();
            p.x = j.x;
            p.y = j.y;
        }
        st_fq(&table[(size_t)w * n + i].x, p.x);
        st_fq(&table[(size_t)w * n + i].y, p.y);
    }
}

// The same table with one field inversion per point instead of one per table entry (380 of the
// 580 multiplications an entry costs above): the doubling chain runs unnormalised, every multiple is
// parked as (x, y) in its table slot and (zz, zzz, running product of the zzz) in `tmp`; the single
// inverse of the total product is then peeled back window by window (Montgomery's trick along the
// chain) and the slots are rewritten in affine form.  tmp holds 3 field elements per entry of
// windows 1 .. nwin-1.  Bit-identical table.
__global__ void __launch_bounds__(128) msm_precompute_batched_kernel(const G1Affine* __restrict__ bases, size_t n, uint32_t c,
                                                                     uint32_t nwin, G1Affine* __restrict__ table,
                                                                     Fq* __restrict__ tmp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = ldg_fq(&bases[i].x);
    p.y = ldg_fq(&bases[i].y);
    st_fq(&table[i].x, p.x);
    st_fq(&table[i].y, p.y);
    if (p.is_identity()) {
        for (uint32_t w = 1; w < nwin; ++w) {
            st_fq(&table[(size_t)w * n + i].x, p.x);
            st_fq(&table[(size_t)w * n + i].y, p.y);
        }
        return;
    }
    G1Xyzz a = G1Xyzz::double_affine(p);
    Fq prod = Fq::one();
    for (uint32_t w = 1; w < nwin; ++w) {
        for (uint32_t k = (w == 1 ? 1u : 0u); k < c; ++k) a = a.dbl();
        Fq* t = tmp + ((size_t)(w - 1) * n + i) * 3;
        st_fq(&table[(size_t)w * n + i].x, a.x);
        st_fq(&table[(size_t)w * n + i].y, a.y);
        st_fq(t, a.zz);
        st_fq(t + 1, a.zzz);
        st_fq(t + 2, prod);                       // product of the zzz of windows 1 .. w-1
        prod = prod * a.zzz;
    }
    Fq inv = prod.inverse();                      // 1 / (zzz_1 ... zzz_{nwin-1})
    for (uint32_t w = nwin - 1; w >= 1; --w) {
        const Fq* t = tmp + ((size_t)(w - 1) * n + i) * 3;
        const Fq zz = ld_fq(t), zzz = ld_fq(t + 1), before = ld_fq(t + 2);
        const Fq zi = inv * before;               // 1 / zzz_w
        inv = inv * zzz;
        const Fq zzi = zi.sqr() * zz.sqr();       // 1 / zz_w = z^-6 z^4
        G1Affine* slot = &table[(size_t)w * n + i];
        const Fq x = ld_fq(&slot->x), y = ld_fq(&slot->y);
        st_fq(&slot->x, (x * zzi).canon());
        st_fq(&slot->y, (y * zi).canon());
    }
}

__global__ void g1_sum_kernel(const G1Jacobian* __restrict__ pts, uint32_t count, G1Jacobian* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t i = 0; i < count; ++i) {
        G1Jacobian j = pts[i];
        if (j.z.is_zero()) continue;
        G1Xyzz p;  // Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3)
        p.x = j.x; p.y = j.y; p.zz = j.z.sqr(); p.zzz = p.zz * j.z;
        acc.add(p);
    }
    *out = acc.to_jacobian();
}

// ------------------------------------------------------------------ host pipeline
// ------------------------------------------------------------- per-stage timing
// Optional CUDA-event bracketing of the pipeline stages (b200zk_msm_profile), used by
// bench.py to time the dominant kernel on the stream it runs on.
enum MsmStage { MSM_ST_HIST = 0, MSM_ST_SCAN, MSM_ST_SCATTER, MSM_ST_SYNC, MSM_ST_ACCUM, MSM_ST_COMBINE,
                MSM_ST_REDUCE, MSM_ST_REDUCE_COMBINE, MSM_ST_FOLD, MSM_ST_END, MSM_ST_COUNT };
struct MsmInfo { uint64_t n; uint32_t c, nwin, npairs, chunk; };
static bool g_msm_profile = false;
static uint32_t g_msm_max_chunk = 128;   // tunable (b200zk_msm_tune)
static bool g_msm_seg_pinned = getenv("B200ZK_MSM_SEG_DIV") || getenv("B200ZK_MSM_MAX_SEGLEN");   // experiments
static uint32_t g_msm_seg_div = getenv("B200ZK_MSM_SEG_DIV") ? (uint32_t)atoi(getenv("B200ZK_MSM_SEG_DIV")) : 256u;   // segments per SM aimed at
static uint32_t g_msm_max_seglen = getenv("B200ZK_MSM_MAX_SEGLEN") ? (uint32_t)atoi(getenv("B200ZK_MSM_MAX_SEGLEN")) : 64;    // measured: 104 -> 64 takes the 242-column reduce from 3.5 to 3.1 ms
static uint32_t g_msm_force_c = 0;
// scatter sub-range bits; B200ZK_MSM_SUB_BITS overrides the automatic choice (experiments only)
// upload pipeline of b200zk_msm_g1_registered (one large host-side commit): number of point
// ranges and the size from which it is used
constexpr size_t MSM_MAX_PARTS = 16;
// measured at 2^24 (scratch/r2_e2e_commit.py; device-resident commit 38.8 ms): 4 equal ranges 43.9 ms; ranges of
// 1/21, 4/21, 16/21 of the points 41.0 ms; (parts, growth) = (3, 3) 41.4, (4, 4) 41.2, (3, 5) 41.7, (4, 2) 42.3, (2, 8) 44.7
constexpr size_t MSM_PIPE_PARTS_DEFAULT = 3;
constexpr size_t MSM_PIPE_MIN_N_DEFAULT = (size_t)1 << 22;
static size_t g_msm_pipe_parts = getenv("B200ZK_MSM_PIPE_PARTS") ? (size_t)atoi(getenv("B200ZK_MSM_PIPE_PARTS")) : MSM_PIPE_PARTS_DEFAULT;
static size_t g_msm_pipe_min_n = getenv("B200ZK_MSM_PIPE_MIN_N") ? (size_t)atoll(getenv("B200ZK_MSM_PIPE_MIN_N")) : MSM_PIPE_MIN_N_DEFAULT;
// ratio of consecutive range lengths: the copy of a range takes ~1/4 of the time its sort + accumulation does, so a
// short first range (little to wait for) followed by ranges growing by that factor keeps the copy ahead of the compute
static double g_msm_pipe_growth = getenv("B200ZK_MSM_PIPE_GROWTH") ? atof(getenv("B200ZK_MSM_PIPE_GROWTH")) : 4.0;
// Unless the schedule is pinned (environment, b200zk_msm_upload_pipeline with parts > 0) it follows the host link: every
// piped commit times its copies and itself, and the next one uses growth = 0.9 x (commit time / copy time), in 3 ranges
// from growth 3 up, 4 below.  One B200 alone copies ~4x faster than it computes (1/21, 4/21, 16/21); eight ranks
// sharing the host's memory copy at 23 GB/s each, ~2x faster, and with growth 4 the last range arrived 8 ms late
// (e2e step at 8 GPUs 73 -> 81 ms).
static bool g_msm_pipe_auto = !getenv("B200ZK_MSM_PIPE_GROWTH") && !getenv("B200ZK_MSM_PIPE_PARTS");
static double g_msm_pipe_ratio = 0.0;       // commit time / copy time of the previous piped commits (0: not measured yet)
static cudaEvent_t g_msm_time_ev[4];        // copies begin / end, commit begin / end
static cudaStream_t g_msm_copy_stream = nullptr;
static cudaEvent_t g_msm_part_ev[MSM_MAX_PARTS + 1];
static void msm_release_pipeline() {          // b200zk_shutdown: the next init may bind another device
    if (!g_msm_copy_stream) return;
    cudaStreamDestroy(g_msm_copy_stream);
    for (auto& e : g_msm_part_ev) cudaEventDestroy(e);
    for (auto& e : g_msm_time_ev) cudaEventDestroy(e);
    g_msm_copy_stream = nullptr;
    g_msm_pipe_ratio = 0.0;
}
static void msm_release_stage_events();
static uint32_t g_msm_chunk_div_small = getenv("B200ZK_MSM_CHUNK_DIV_SMALL") ? (uint32_t)atoi(getenv("B200ZK_MSM_CHUNK_DIV_SMALL")) : 512u;
static uint32_t g_msm_chunk_div = getenv("B200ZK_MSM_CHUNK_DIV") ? (uint32_t)atoi(getenv("B200ZK_MSM_CHUNK_DIV")) : 2048u;
static uint32_t g_msm_force_sub = getenv("B200ZK_MSM_SUB_BITS") ? (uint32_t)atoi(getenv("B200ZK_MSM_SUB_BITS")) : 0xffffffffu;
// stage events are kept per stream (a profiled MSM on one stream does not disturb another's);
// b200zk_msm_last_stages reports the most recent profiled call
struct MsmStageEvents {
    cudaEvent_t ev[MSM_ST_COUNT];
    bool valid[MSM_ST_COUNT];
    MsmInfo info;
    const MsmRun* d_run = nullptr;     // device plan of that call (pair count, chunk length)
};
static std::map<cudaStream_t, MsmStageEvents*> g_msm_stage_events;
static MsmStageEvents* g_msm_last_profiled = nullptr;
static cudaStream_t g_msm_ev_stream = nullptr;
static void msm_release_stage_events() {
    for (auto& kv : g_msm_stage_events) {
        for (auto& e : kv.second->ev) cudaEventDestroy(e);
        delete kv.second;
    }
    g_msm_stage_events.clear();
    g_msm_last_profiled = nullptr;
}

struct StageTimer {
    cudaStream_t s;
    bool on;
    MsmStageEvents* E = nullptr;
    StageTimer(Context&, cudaStream_t st) : s(st), on(g_msm_profile) {
        if (!on) return;
        auto it = g_msm_stage_events.find(st);
        if (it == g_msm_stage_events.end()) {
            E = new MsmStageEvents();
            for (int i = 0; i < MSM_ST_COUNT; ++i) ZK_CUDA(cudaEventCreate(&E->ev[i]));
            g_msm_stage_events[st] = E;
        } else E = it->second;
        for (int i = 0; i < MSM_ST_COUNT; ++i) E->valid[i] = false;
        g_msm_ev_stream = st;
        g_msm_last_profiled = E;
    }
    void mark(int stage) {
        if (!on) return;
        ZK_CUDA(cudaEventRecord(E->ev[stage], s));
        E->valid[stage] = true;
    }
};

// minimise 10*n*W (mixed adds) + 30*G*2^(c-1) (two full adds per bucket + slack), where G
// is the number of bucket sets: W without a precomputed table, 1 with one.
static uint32_t choose_window(size_t n, bool shared_buckets) {
    double best = 1e300;
    uint32_t bc = 4;
    for (uint32_t c = 4; c <= 23; ++c) {
        const double W = (255 + c - 1) / c;
        const double cost = 10.0 * (double)n * W + 30.0 * (shared_buckets ? 1.0 : W) * (double)(1u << (c - 1));
        if (cost < best) { best = cost; bc = c; }
    }
    // up to 2^18 points one commit is latency-bound: 2^15 buckets keep the bucket reduction in its 4-lane
    // cooperative form (<= 2^14 segments), worth more than the 1/16 fewer additions of one more bit
    // (2^17 points: 1.01 ms at c = 17, 0.89 ms at c = 16)
    if (shared_buckets && n <= ((size_t)1 << 18) && bc > 16) bc = 16;
    return bc;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct MsmPre {
    const G1Affine* table;
    size_t n_reg;
    uint32_t c, nwin;
};

// Launch the keyed-reduction levels on (keys, pts)[count] until everything has been
// added into `buckets`.  Small inputs use short chunks: the cost of a level is the latency
// of L dependent point additions, not throughput.
static bool g_msm_quad = getenv("B200ZK_MSM_QUAD") ? atoi(getenv("B200ZK_MSM_QUAD")) != 0 : true;
static uint32_t g_msm_finish_max = getenv("B200ZK_MSM_FINISH_MAX") ? (uint32_t)atoi(getenv("B200ZK_MSM_FINISH_MAX")) : 48u;
static size_t g_msm_quad_reduce_max = getenv("B200ZK_MSM_QUAD_REDUCE_MAX") ? (size_t)atoll(getenv("B200ZK_MSM_QUAD_REDUCE_MAX")) : ((size_t)1 << 14);
static bool g_msm_psort = getenv("B200ZK_MSM_PSORT") ? atoi(getenv("B200ZK_MSM_PSORT")) != 0 : true;
static uint32_t g_msm_psort_shift = getenv("B200ZK_MSM_PSORT_SHIFT") ? std::min<uint32_t>((uint32_t)atoi(getenv("B200ZK_MSM_PSORT_SHIFT")), PSORT_MAX_SHIFT) : 11u;
static size_t g_msm_psort_min_pairs = getenv("B200ZK_MSM_PSORT_MIN_PAIRS") ? (size_t)atoll(getenv("B200ZK_MSM_PSORT_MIN_PAIRS")) : ((size_t)1 << 22);
static uint32_t g_msm_finish_max_keys = getenv("B200ZK_MSM_FINISH_MAX_KEYS") ? (uint32_t)atoi(getenv("B200ZK_MSM_FINISH_MAX_KEYS")) : (1u << 18);
static uint32_t g_msm_finish_max_threads = getenv("B200ZK_MSM_FINISH_MAX_THREADS") ? (uint32_t)atoi(getenv("B200ZK_MSM_FINISH_MAX_THREADS")) : (1u << 19);
static uint32_t g_msm_quad_max = getenv("B200ZK_MSM_QUAD_MAX") ? (uint32_t)atoi(getenv("B200ZK_MSM_QUAD_MAX")) : (1u << 17);

static void run_combine_levels(Context& c, uint32_t* keysA, G1Xyzz* ptsA, uint32_t* keysB, G1Xyzz* ptsB,
                               uint32_t count, uint32_t* counts, G1Xyzz* buckets, cudaStream_t s,
                               const uint32_t* enabled = nullptr) {
    // `count` bounds the entries of level 0; the exact counts are in counts[level] (device)
    uint32_t level = 0;
    while (count > 0) {
        ZK_REQUIRE(level + 1 < (uint32_t)MSM_MAX_LEVELS, "too many keyed-reduction levels");
        if (enabled && count > 1024u) {
            // behind the direct finish these levels only run for skewed scalars (a bucket longer than the
            // finish takes); otherwise each is an empty launch, so shrink fast: 16x, then 32x per level
            const uint32_t L = 16u;
            const uint32_t nthreads = (count + L - 1) / L;
            msm_combine_kernel<<<(nthreads + 127) / 128, 128, 0, s>>>(keysA, ptsA, counts, level, L, buckets, keysB, ptsB, enabled);
            ZK_LAUNCH_CHECK();
            count = nthreads;
        } else if (enabled) {
            const uint32_t nwarps = (count + 31) / 32;
            msm_combine_warp_kernel<<<(nwarps * 32 + 127) / 128, 128, 0, s>>>(keysA, ptsA, counts, level, buckets, keysB, ptsB, enabled);
            ZK_LAUNCH_CHECK();
            if (nwarps == 1) break;
            count = nwarps;
        } else if (count > (g_msm_quad ? g_msm_quad_max : 32768u)) {
            // throughput regime: sequential chunks, one addition per entry
            const uint32_t L = count > (uint32_t)c.sm_count * 4096u ? 16u : 4u;
            const uint32_t nthreads = (count + L - 1) / L;
            msm_combine_kernel<<<(nthreads + 127) / 128, 128, 0, s>>>(keysA, ptsA, counts, level, L, buckets, keysB, ptsB, enabled);
            ZK_LAUNCH_CHECK();
            count = nthreads;
        } else if (g_msm_quad) {
            // latency regime: one entry per quad, 8x per level, additions four multiplications deep
            const uint32_t nwarps = (count + 7) / 8;
            msm_combine_quad_kernel<<<(nwarps * 32 + 127) / 128, 128, 0, s>>>(keysA, ptsA, counts, level, buckets, keysB, ptsB, enabled);
            ZK_LAUNCH_CHECK();
            if (nwarps == 1) break;
            count = nwarps;
        } else {
            const uint32_t nwarps = (count + 31) / 32;
            msm_combine_warp_kernel<<<(nwarps * 32 + 127) / 128, 128, 0, s>>>(keysA, ptsA, counts, level, buckets, keysB, ptsB, enabled);
            ZK_LAUNCH_CHECK();
            if (nwarps == 1) break;
            count = nwarps;
        }
        ++level;
        std::swap(keysA, keysB);
        std::swap(ptsA, ptsB);
    }
}

// Sum `per_group` consecutive partials of each group: grid (nsplit, groups), 256 threads;
// thread j adds its strided share, then a shared-memory tree.  out[group * nsplit + split].
__global__ void __launch_bounds__(256) msm_group_sum_kernel(const G1Xyzz* __restrict__ pts, uint32_t per_group,
                                                            uint32_t nsplit, G1Xyzz* __restrict__ out) {
    extern __shared__ uint4 gs_smem[];
    G1Xyzz* sh = reinterpret_cast<G1Xyzz*>(gs_smem);
    const uint32_t group = blockIdx.y, split = blockIdx.x;
    const uint32_t span = (per_group + nsplit - 1) / nsplit;
    const uint32_t lo = split * span, hi = min(lo + span, per_group);
    const G1Xyzz* base = pts + (size_t)group * per_group;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) {
        G1Xyzz p = ld_xyzz(base + i);
        acc.add(p);
    }
    st_xyzz(sh + threadIdx.x, acc);
    __syncthreads();
    for (uint32_t stride = 128; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) {
            G1Xyzz a = ld_xyzz(sh + threadIdx.x);
            G1Xyzz b = ld_xyzz(sh + threadIdx.x + stride);
            a.add(b);
            st_xyzz(sh + threadIdx.x, a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(out + (size_t)group * nsplit + split, ld_xyzz(sh));
}

// `count` MSMs over the same n bases: column j uses scalars d_scalars + j * scalar_stride and
// writes d_out[j].  All inputs device-resident.  `pre` (optional) is a per-window table.
// One MSM may be fed in consecutive point ranges that share the bucket set (a host-side commit
// whose scalars are still arriving over PCIe): `first` clears the buckets, later parts add into
// them, `last` runs the bucket reduction.  index_offset = first point of the range.
struct MsmPart {
    bool first = true, last = true;
    uint32_t index_offset = 0;
    // length of the longest range of this MSM: every range sizes the work arena for it, so that the arena — whose
    // head holds the buckets the ranges share — never has to grow (and move) between two ranges
    size_t longest = 0;
};

static void msm_device(Context& c, const Fr* d_scalars, size_t scalar_stride, size_t count,
                       const G1Affine* d_bases, size_t n, const MsmPre* pre, G1Jacobian* d_out, cudaStream_t s,
                       const MsmPart& part = MsmPart()) {
    ZK_REQUIRE(n < ((size_t)1 << 28), "MSM size must be below 2^28 points");
    if (count == 0) return;
    if (n == 0) {
        G1Jacobian id;
        id.x = Fq::zero(); id.y = Fq::one(); id.z = Fq::zero();
        for (size_t j = 0; j < count; ++j)
            ZK_CUDA(cudaMemcpyAsync(d_out + j, &id, sizeof id, cudaMemcpyHostToDevice, s));
        ZK_CUDA(cudaStreamSynchronize(s));
        return;
    }
    const uint32_t cbits = pre ? pre->c : (g_msm_force_c ? g_msm_force_c : choose_window(n, false));
    const uint32_t nwin = (255 + cbits - 1) / cbits;
    const uint32_t key_windows = pre ? 1u : nwin;
    const uint32_t nb = 1u << (cbits - 1);
    const size_t groups = count * key_windows;            // bucket sets
    const size_t nkeys_sz = groups * nb;
    const size_t max_pairs = n * nwin * count;
    ZK_REQUIRE(max_pairs < ((size_t)1 << 32) && nkeys_sz < ((size_t)1 << 31), "MSM batch too large for 32-bit offsets");
    ZK_REQUIRE(count * nwin <= 65535, "MSM batch has too many (column, window) pairs");
    if (pre) ZK_REQUIRE(pre->n_reg * (size_t)nwin < ((size_t)1 << 31), "precomputed table too large for 31-bit indices");
    const uint32_t nkeys = (uint32_t)nkeys_sz;

    // ---- carve the sort arena (sizes known up front)
    const uint32_t scan_blocks = (uint32_t)((nkeys + SCAN_BLOCK - 1) / SCAN_BLOCK);
    // segment length of the bucket reduction: enough segments to fill the machine, short
    // enough that 2 * seglen dependent additions stay cheap
    // Measured on B200 (scratch/r2_seg_sweep.sh, batches of 2^15-point columns, 2^14 buckets each; reduce + sum of the
    // segment partials): below 2^20 buckets ~96 segments per SM — few enough for the 4-lane cooperative kernel, at most
    // 128 buckets each — beat 256 per SM (8 columns 0.54 -> 0.37 ms, 16: 0.65 -> 0.49, 32: 0.84 -> 0.74); from 2^20
    // buckets up one lane per segment of <= 64 buckets wins (64 columns 1.18 against 1.27 ms, 2^21 buckets 1.92 / 2.02)
    const bool mid = nkeys < (1u << 20) && !g_msm_seg_pinned;
    const uint32_t seg_div = mid ? 96u : g_msm_seg_div;
    const uint32_t max_seglen = mid ? std::max<uint32_t>(g_msm_max_seglen, 128u) : g_msm_max_seglen;
    uint32_t seglen = (uint32_t)std::min<uint64_t>(max_seglen, std::max<uint64_t>(4, nkeys / ((uint64_t)c.sm_count * seg_div)));
    if (seglen > nb) seglen = nb;
    const uint32_t segs_per_group = (nb + seglen - 1) / seglen;
    const size_t red_entries = (size_t)segs_per_group * groups;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_buckets = carve((size_t)nkeys * sizeof(G1Xyzz));   // first: same place for every part of an MSM
    const size_t o_win = carve(groups * sizeof(G1Xyzz));
    const size_t o_hist = carve(((size_t)nkeys + 1) * 4);
    const size_t o_start = carve(((size_t)nkeys + 1) * 4);
    const size_t o_cursor = carve(((size_t)nkeys + 1) * 4);
    const size_t o_bsum = carve(((size_t)scan_blocks + 1) * 4);
    const size_t o_total = carve(256);
    const size_t o_run = carve(sizeof(MsmRun));
    const size_t pairs_cap = std::max(n, part.longest) * nwin * count;
    const size_t o_sorted = carve(pairs_cap * 4);
    const size_t o_digits = carve(pairs_cap * 4);
    // two-level partition sort (large inputs): 8-byte records + the partition tables
    const bool psort_fits = g_msm_psort && nkeys <= ((size_t)PSORT_MAX_PARTS << PSORT_MAX_SHIFT);
    const bool psort = psort_fits && max_pairs >= g_msm_psort_min_pairs;
    const bool psort_cap = psort_fits && pairs_cap >= g_msm_psort_min_pairs;
    const size_t o_rec = carve(psort_cap ? pairs_cap * 8 : 0);
    const size_t o_ptab = carve(psort_cap ? (size_t)(4 * (PSORT_MAX_PARTS + 1)) * 4 : 0);
    char* base = (char*)c.scratch(s).msm_work.get(off);
    uint32_t* hist = (uint32_t*)(base + o_hist);
    uint32_t* start = (uint32_t*)(base + o_start);
    uint32_t* cursor = (uint32_t*)(base + o_cursor);
    uint32_t* bsum = (uint32_t*)(base + o_bsum);
    uint32_t* total = (uint32_t*)(base + o_total);
    MsmRun* run = (MsmRun*)(base + o_run);
    uint32_t* sorted = (uint32_t*)(base + o_sorted);
    int32_t* digits = (int32_t*)(base + o_digits);
    G1Xyzz* buckets = (G1Xyzz*)(base + o_buckets);
    G1Xyzz* win = (G1Xyzz*)(base + o_win);

    StageTimer T(c, s);
    // ---- 1-3: sort (bucket, point) pairs
    ZK_CUDA(cudaMemsetAsync(hist, 0, ((size_t)nkeys + 1) * 4, s));
    if (part.first) {
        ZK_CUDA(cudaMemsetAsync(buckets, 0, (size_t)nkeys * sizeof(G1Xyzz), s));
        ZK_CUDA(cudaMemsetAsync(win, 0, groups * sizeof(G1Xyzz), s));
    }
    const unsigned sblocks = (unsigned)((n + 255) / 256);
    T.mark(MSM_ST_HIST);
    PsortArgs PA{};
    uint32_t* ptab = (uint32_t*)(base + o_ptab);
    uint32_t* part_count = ptab, *part_start = ptab + (PSORT_MAX_PARTS + 1), *part_cursor = ptab + 2 * (PSORT_MAX_PARTS + 1),
            *tile_prefix = ptab + 3 * (PSORT_MAX_PARTS + 1);
    uint2* rec = (uint2*)(base + o_rec);
    const unsigned ptiles_a = (unsigned)((max_pairs + PSORT_TILE_A - 1) / PSORT_TILE_A);
    const unsigned ptiles_b = (unsigned)((max_pairs + PSORT_TILE_B - 1) / PSORT_TILE_B);
    if (psort) {
        PA.c = cbits; PA.nwin = nwin; PA.key_windows = key_windows;
        // as few partitions as the per-partition bucket count (<= 2^11) allows: the partition cursors are the one
        // set of counters every tile of pass A contends for (2^21 keys: 1024 partitions of 2048 buckets)
        PA.shift = g_msm_psort_shift;
        while (PA.shift > 0 && ((size_t)1 << PA.shift) > (size_t)nkeys) --PA.shift;
        while ((((size_t)nkeys + ((size_t)1 << PA.shift) - 1) >> PA.shift) > PSORT_MAX_PARTS) ++PA.shift;
        PA.nparts = (uint32_t)(((size_t)nkeys + ((size_t)1 << PA.shift) - 1) >> PA.shift);
        PA.table_stride = pre ? (uint32_t)pre->n_reg : 0u;
        PA.index_offset = part.index_offset;
        ZK_CUDA(cudaMemsetAsync(part_count, 0, (PSORT_MAX_PARTS + 1) * 4, s));
        msm_digits_count_kernel<<<dim3((unsigned)((n + 2047) / 2048), (unsigned)count), 256, 0, s>>>(d_scalars, scalar_stride, n, PA,
                                                                                                       part_count, digits);
        ZK_LAUNCH_CHECK();
        msm_partition_scan_kernel<<<1, 1024, 0, s>>>(part_count, PA.nparts, part_start, part_cursor, tile_prefix);
        ZK_LAUNCH_CHECK();
        constexpr int smemA = PSORT_TILE_A * 8 + 2 * PSORT_MAX_PARTS * 4 + (PSORT_THREADS_A + 4) * 4;
        constexpr int smemB = PSORT_TILE_B * 8 + 2 * (1 << PSORT_MAX_SHIFT) * 4 + (PSORT_THREADS_B + 4) * 4;
        static int psort_configured_device = -1;
        if (psort_configured_device != c.device) {
            ZK_CUDA(cudaFuncSetAttribute(msm_partition_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemA));
            ZK_CUDA(cudaFuncSetAttribute(msm_bucket_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB));
            psort_configured_device = c.device;
        }
        msm_partition_scatter_kernel<<<ptiles_a, PSORT_THREADS_A, smemA, s>>>(digits, n, (uint64_t)max_pairs, PA, part_cursor, rec);
        ZK_LAUNCH_CHECK();
        msm_bucket_count_kernel<<<ptiles_b + PA.nparts, PSORT_THREADS_B, 0, s>>>(rec, PA, tile_prefix, part_start, hist);
        ZK_LAUNCH_CHECK();
    } else {
        msm_hist_kernel<<<dim3(sblocks, (unsigned)count), 256, 0, s>>>(d_scalars, scalar_stride, n, cbits, nwin,
                                                                        key_windows, hist, digits);
        ZK_LAUNCH_CHECK();
    }
    T.mark(MSM_ST_SCAN);
    scan_local_kernel<<<scan_blocks, SCAN_THREADS, 0, s>>>(hist, start, bsum, nkeys);
    ZK_LAUNCH_CHECK();
    scan_sums_kernel<<<1, SCAN_THREADS, 0, s>>>(bsum, scan_blocks, total);
    ZK_LAUNCH_CHECK();
    // chunk length: keep >= ~2048 threads per SM queued, between 4 and 128 pairs each (longer
    // chunks mean fewer open runs handed to the keyed-reduction levels; measured on B200 against
    // 4096 / 1024 / 512: 1.79 -> 1.59 ms at 2^18, 4.27 -> 3.99 ms at 2^20, unchanged from 2^23 up).
    // Chosen by the device from the pair count it has just produced (MsmRun).
    // A single small commit (latency regime, finished directly by quads) prefers ~4x longer chunks: half as many
    // open partials per bucket for the finish (2^15 points: 0.441 -> 0.416 ms per commit at 512 against 2048).
    const bool small_regime = g_msm_quad && max_pairs <= ((size_t)1 << 22) && nkeys <= g_msm_finish_max_keys;
    const uint32_t chunk_unit = (uint32_t)c.sm_count * (small_regime ? std::min<uint32_t>(g_msm_chunk_div, g_msm_chunk_div_small) : g_msm_chunk_div);
    scan_add_kernel<<<scan_blocks, SCAN_THREADS, 0, s>>>(start, cursor, bsum, nkeys, total, chunk_unit, g_msm_max_chunk, run);
    ZK_LAUNCH_CHECK();
    T.mark(MSM_ST_SCATTER);
    // bucket sub-ranges: measured on B200 at k = 24, one split helps the shared-bucket (table)
    // layout (6.4 -> 5.5 ms) and none helps the per-window layout
    uint32_t sub_bits = (pre && n * 4 > ((size_t)16 << 20) && cbits > 2 && count * nwin * 2 <= 65535) ? 1u : 0u;
    if (g_msm_force_sub != 0xffffffffu) sub_bits = std::min<uint32_t>(g_msm_force_sub, cbits - 1);
    if (psort) {
        msm_bucket_scatter_kernel<<<ptiles_b + PA.nparts, PSORT_THREADS_B, PSORT_TILE_B * 8 + 2 * (1 << PSORT_MAX_SHIFT) * 4 + (PSORT_THREADS_B + 4) * 4, s>>>(
            rec, PA, tile_prefix, part_start, cursor, sorted);
    } else {
        msm_scatter_kernel<<<dim3(sblocks, (unsigned)((count * nwin) << sub_bits)), 256, 0, s>>>(
            digits, n, cbits, nwin, key_windows, pre ? (uint32_t)pre->n_reg : 0u, sub_bits, part.index_offset, cursor, sorted);
    }
    ZK_LAUNCH_CHECK();
    T.mark(MSM_ST_SYNC);
    // No host round trip for the pair count: the accumulation is launched for the largest thread count any
    // pair count <= max_pairs can plan (np / L(np) with L = clamp(np / chunk_unit, 4, max_chunk)).
    const uint64_t bound_short = (max_pairs + 3) / 4;                                   // L = 4
    const uint64_t bound_long = std::max<uint64_t>((uint64_t)chunk_unit * 5 / 4 + 2,   // 4 < L < max: np / floor(np / unit)
                                                   max_pairs / g_msm_max_chunk + 1);     // L = max_chunk
    const uint32_t nthreads0 = (uint32_t)std::max<uint64_t>(1, std::min(bound_short, bound_long));
    // ---- carve the carry arena for that bound
    const size_t carry_cap = std::max<size_t>(std::max<size_t>(nthreads0, red_entries), 64);
    off = 0;
    const size_t o_keyA = carve(carry_cap * 4);
    const size_t o_ptA = carve(carry_cap * sizeof(G1Xyzz));
    const size_t capB = std::max<size_t>(carry_cap / 4 + 64, groups * 64);
    const size_t o_keyB = carve(capB * 4);
    const size_t o_ptB = carve(capB * sizeof(G1Xyzz));
    char* cbase = (char*)c.scratch(s).msm_carry.get(off);
    uint32_t* keyA = (uint32_t*)(cbase + o_keyA);
    G1Xyzz* ptA = (G1Xyzz*)(cbase + o_ptA);
    uint32_t* keyB = (uint32_t*)(cbase + o_keyB);
    G1Xyzz* ptB = (G1Xyzz*)(cbase + o_ptB);
    if (T.on) {
        T.E->info.n = n * count; T.E->info.c = cbits; T.E->info.nwin = nwin; T.E->info.npairs = 0; T.E->info.chunk = 0;
        T.E->d_run = run;
    }

    // ---- 4: accumulate
    T.mark(MSM_ST_ACCUM);
    if (part.first)
        msm_accum_kernel<false><<<(nthreads0 + 127) / 128, 128, 0, s>>>(pre ? pre->table : d_bases, sorted, start, nkeys, run,
                                                                        buckets, keyA, ptA);
    else
        msm_accum_kernel<true><<<(nthreads0 + 127) / 128, 128, 0, s>>>(pre ? pre->table : d_bases, sorted, start, nkeys, run,
                                                                       buckets, keyA, ptA);
    ZK_LAUNCH_CHECK();
    T.mark(MSM_ST_COMBINE);
    if (g_msm_quad && nthreads0 <= g_msm_finish_max_threads && nkeys <= g_msm_finish_max_keys) {
        // latency regime (a single small commit): every bucket's open partials by one quad; the keyed levels below
        // only run when a bucket was too long for it.  (With 2^21 buckets — a 2^24 commit, or one point range of its
        // upload pipeline — the levels are cheaper: 0.64 ms against 1.56 ms.)
        msm_finish_quad_kernel<<<(nkeys * 4u + 127u) / 128u, 128, 0, s>>>(start, nkeys, run, g_msm_finish_max, buckets, ptA);
        ZK_LAUNCH_CHECK();
        run_combine_levels(c, keyA, ptA, keyB, ptB, nthreads0, run->level_count, buckets, s, &run->any_long);
    } else {
        run_combine_levels(c, keyA, ptA, keyB, ptB, nthreads0, run->level_count, buckets, s);
    }

    if (!part.last) {
        T.mark(MSM_ST_END);
        return;
    }
    // ---- 5: per-bucket-set running sums, then keyed reduction with key = bucket set
    T.mark(MSM_ST_REDUCE);
    // quads for the latency regime; a batch of columns or a 2^24-point commit has enough segments to keep the
    // multipliers busy with one lane per segment, which does ~1.5x less work
    if (g_msm_quad && red_entries <= g_msm_quad_reduce_max) {
        uint32_t lo_bits = 0;
        while (lo_bits < 32 && (((uint64_t)(segs_per_group - 1) * seglen) >> lo_bits) != 0) ++lo_bits;
        const uint32_t nthreads = (uint32_t)red_entries * 4;
        msm_reduce_quad_kernel<<<(nthreads + 127) / 128, 128, 0, s>>>(buckets, nb, seglen, segs_per_group, (uint32_t)groups,
                                                                      lo_bits, run->level_count2, keyA, ptA);
        ZK_LAUNCH_CHECK();
        T.mark(MSM_ST_REDUCE_COMBINE);
        // per bucket set: the segment partials (key = bucket set, sorted) by the same keyed reduction, into `win`
        run_combine_levels(c, keyA, ptA, keyB, ptB, (uint32_t)red_entries, run->level_count2, win, s);
    } else {
        const uint32_t nthreads = (uint32_t)red_entries;
        msm_reduce_kernel<<<(nthreads + 127) / 128, 128, 0, s>>>(buckets, nb, seglen, segs_per_group, (uint32_t)groups,
                                                                 keyA, ptA);
        ZK_LAUNCH_CHECK();
        T.mark(MSM_ST_REDUCE_COMBINE);
        // per bucket set: sum its segment partials (block tree, two stages when long)
        const uint32_t nsplit = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, segs_per_group / 2048));
        const int smem = 256 * sizeof(G1Xyzz);
        if (nsplit == 1) {
            msm_group_sum_kernel<<<dim3(1, (unsigned)groups), 256, smem, s>>>(ptA, segs_per_group, 1, win);
            ZK_LAUNCH_CHECK();
        } else {
            msm_group_sum_kernel<<<dim3(nsplit, (unsigned)groups), 256, smem, s>>>(ptA, segs_per_group, nsplit, ptB);
            ZK_LAUNCH_CHECK();
            msm_group_sum_kernel<<<dim3(1, (unsigned)groups), 256, smem, s>>>(ptB, nsplit, 1, win);
            ZK_LAUNCH_CHECK();
        }
    }
    T.mark(MSM_ST_FOLD);
    // ---- 6: fold each column's windows (a single one with a precomputed table)
    msm_fold_kernel<<<(unsigned)((count * 4 + 31) / 32), 32, 0, s>>>(win, key_windows, cbits, (uint32_t)count, d_out);
    ZK_LAUNCH_CHECK();
    T.mark(MSM_ST_END);
}

static void copy_point_out(Context& c, const G1Jacobian* d, uint64_t* out, cudaStream_t s) {
    ZK_CUDA(cudaMemcpyAsync(out, d, sizeof(G1Jacobian), cudaMemcpyDeviceToHost, s));
    ZK_CUDA(cudaStreamSynchronize(s));
    (void)c;
}

}  // namespace zk

using namespace zk;

extern "C" {

int b200zk_msm_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_xyz[12]) {
    return guarded([&] {
        ZK_REQUIRE(out_xyz && (n == 0 || (scalars && bases)), "null argument");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        Fr* ds = (Fr*)c.scratch(s).msm_scalars.get(std::max<size_t>(n, 1) * sizeof(Fr));
        G1Affine* db = (G1Affine*)c.scratch(s).msm_bases.get(std::max<size_t>(n, 1) * sizeof(G1Affine));
        G1Jacobian* dout = (G1Jacobian*)c.scratch(s).misc.get(sizeof(G1Jacobian));
        if (n) {
            ZK_CUDA(cudaMemcpyAsync(ds, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
            ZK_CUDA(cudaMemcpyAsync(db, bases, n * sizeof(G1Affine), cudaMemcpyHostToDevice, s));
        }
        msm_device(c, ds, n, 1, db, n, nullptr, dout, s);
        copy_point_out(c, dout, out_xyz, s);
    });
}

static void register_bases(const uint64_t* bases, size_t n, int precompute, uint64_t* handle_out) {
    ZK_REQUIRE(handle_out && (n == 0 || bases), "null argument");
    ensure_init();
    Context& c = ctx();
    BaseTable* t = new BaseTable();
    t->n = n;
    if (n) {
        ZK_CUDA(cudaMalloc(&t->d, n * sizeof(G1Affine)));
        ZK_CUDA(cudaMemcpy(t->d, bases, n * sizeof(G1Affine), cudaMemcpyHostToDevice));
    }
    if (n && precompute) {
        // precompute = 1: window size from the cost model; >= 4: that many bits (experiments, small-n tuning)
        t->c = precompute >= 4 ? (uint32_t)std::min(precompute, 23) : choose_window(n, true);
        t->nwin = (255 + t->c - 1) / t->c;
        const size_t bytes = n * (size_t)t->nwin * sizeof(G1Affine);
        size_t free_b = 0, total_b = 0;
        ZK_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (n * (size_t)t->nwin < ((size_t)1 << 31) && bytes < free_b / 2) {
            ZK_CUDA(cudaMalloc(&t->table, bytes));
            // one inversion per point needs 96 B of scratch per table entry; fall back to one inversion
            // per entry when that does not fit next to the table
            Fq* tmp = nullptr;
            const size_t tmp_bytes = n * (size_t)(t->nwin - 1) * 3 * sizeof(Fq);
            if (t->nwin > 1 && cudaMalloc(&tmp, tmp_bytes) != cudaSuccess) {
                cudaGetLastError();
                tmp = nullptr;
            }
            if (tmp)
                msm_precompute_batched_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c.stream>>>(t->d, n, t->c, t->nwin, t->table, tmp);
            else
                msm_precompute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c.stream>>>(t->d, n, t->c, t->nwin, t->table);
            ZK_LAUNCH_CHECK();
            ZK_CUDA(cudaStreamSynchronize(c.stream));
            if (tmp) cudaFree(tmp);
        }  // otherwise: plain registered bases (generic pipeline)
    }
    const uint64_t h = c.next_handle++;
    c.bases[h] = t;
    *handle_out = h;
}

int b200zk_bases_register(const uint64_t* bases, size_t n, uint64_t* handle_out) {
    return guarded([&] { register_bases(bases, n, 1, handle_out); });
}

int b200zk_bases_register_ex(const uint64_t* bases, size_t n, int precompute_windows, uint64_t* handle_out) {
    return guarded([&] { register_bases(bases, n, precompute_windows, handle_out); });
}

int b200zk_bases_evict(uint64_t handle) {
    return guarded([&] {
        Context& c = ctx();
        auto it = c.bases.find(handle);
        ZK_REQUIRE(it != c.bases.end(), "unknown bases handle");
        if (c.ready) {
            cudaSetDevice(c.device);
            cudaStreamSynchronize(c.stream);
        }
        cudaFree(it->second->d);
        if (it->second->table) cudaFree(it->second->table);
        delete it->second;
        c.bases.erase(it);
    });
}

// Point ranges of the upload pipeline: up to `parts` consecutive ranges of [0, n) whose lengths grow by `growth` from
// one to the next (1: equal), every boundary but the last a multiple of 256 points.  begin_of[0 .. count] on return.
static size_t msm_upload_ranges(size_t n, size_t parts, double growth, size_t* begin_of) {
    parts = std::max<size_t>(1, std::min<size_t>(parts, MSM_MAX_PARTS));
    const double g = std::max(1.0, growth);
    double sum = 0, w = 1;
    for (size_t p = 0; p < parts; ++p, w *= g) sum += w;
    size_t count = 0, b = 0;
    w = 1;
    for (size_t p = 0; p < parts && b < n; ++p, w *= g) {
        size_t len = p + 1 == parts ? n - b : align_up((size_t)((double)n * (w / sum)) + 1, 256);
        len = std::min(len, n - b);
        begin_of[count++] = b;
        b += len;
    }
    begin_of[count] = n;
    return count;
}

static void msm_registered(uint64_t handle, const uint64_t* scalars, size_t stride, size_t count, size_t n,
                           uint64_t* out_xyz) {
    ZK_REQUIRE(out_xyz && (n == 0 || scalars), "null argument");
    ZK_REQUIRE(count <= 1 || stride >= n, "batch stride smaller than the column");
    ensure_init();
    Context& c = ctx();
    auto it = c.bases.find(handle);
    ZK_REQUIRE(it != c.bases.end(), "unknown bases handle");
    BaseTable* t = it->second;
    ZK_REQUIRE(n <= t->n, "more scalars than registered bases");
    if (count == 0) return;
    cudaStream_t s = c.stream;
    // a single commit of a polynomial the library already mirrors (b200zk_mirror_enable) needs no upload
    Fr* ds = nullptr;
    bool resident = false;
    if (count == 1 && n && c.mirrors.enabled) {
        ds = (Fr*)c.mirrors.find(scalars, n);
        resident = ds != nullptr;
        if (!ds) ds = (Fr*)c.mirrors.insert(scalars, n);
    } else {
        ds = (Fr*)c.scratch(s).msm_scalars.get(std::max<size_t>(n * count, 1) * sizeof(Fr));
    }
    G1Jacobian* dout = (G1Jacobian*)c.scratch(s).misc.get(count * sizeof(G1Jacobian));
    MsmPre pre{t->table, t->n, t->c, t->nwin};
    // One large commit with a window table: feed it in point ranges that share the bucket set,
    // so the upload of range p+1 runs under the sort + accumulation of range p.
    size_t parts = (count == 1 && !resident && t->table && n >= g_msm_pipe_min_n) ? std::min<size_t>(g_msm_pipe_parts, MSM_MAX_PARTS) : 1;
    if (parts > 1) {
        if (!g_msm_copy_stream) {
            ZK_CUDA(cudaStreamCreateWithFlags(&g_msm_copy_stream, cudaStreamNonBlocking));
            for (auto& e : g_msm_part_ev) ZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (auto& e : g_msm_time_ev) ZK_CUDA(cudaEventCreate(&e));
        }
        double growth = std::max(1.0, g_msm_pipe_growth);
        if (g_msm_pipe_auto && g_msm_pipe_ratio > 0.0) {
            growth = std::min(4.0, std::max(1.0, 0.9 * g_msm_pipe_ratio));
            parts = growth >= 3.0 ? 3 : 4;
        }
        // range lengths in geometric progression (growth 1: equal ranges), multiples of 256 points
        size_t begin_of[MSM_MAX_PARTS + 1];
        const size_t nparts = msm_upload_ranges(n, parts, growth, begin_of);
        size_t longest = 0;
        for (size_t p = 0; p < nparts; ++p) longest = std::max(longest, begin_of[p + 1] - begin_of[p]);
        ZK_CUDA(cudaEventRecord(g_msm_part_ev[MSM_MAX_PARTS], s));   // earlier users of the staging buffer
        ZK_CUDA(cudaStreamWaitEvent(g_msm_copy_stream, g_msm_part_ev[MSM_MAX_PARTS], 0));
        ZK_CUDA(cudaEventRecord(g_msm_time_ev[0], g_msm_copy_stream));
        ZK_CUDA(cudaEventRecord(g_msm_time_ev[2], s));
        for (size_t p = 0; p < nparts; ++p) {
            const size_t b = begin_of[p], len = begin_of[p + 1] - b;
            ZK_CUDA(cudaMemcpyAsync(ds + b, scalars + 4 * b, len * sizeof(Fr), cudaMemcpyHostToDevice, g_msm_copy_stream));
            ZK_CUDA(cudaEventRecord(g_msm_part_ev[p], g_msm_copy_stream));
        }
        ZK_CUDA(cudaEventRecord(g_msm_time_ev[1], g_msm_copy_stream));
        for (size_t p = 0; p < nparts; ++p) {
            const size_t b = begin_of[p], len = begin_of[p + 1] - b;
            ZK_CUDA(cudaStreamWaitEvent(s, g_msm_part_ev[p], 0));
            MsmPart part;
            part.first = p == 0; part.last = p + 1 == nparts; part.index_offset = (uint32_t)b;
            part.longest = longest;
            msm_device(c, ds + b, len, 1, t->d, len, &pre, dout, s, part);
        }
        ZK_CUDA(cudaMemcpyAsync(out_xyz, dout, sizeof(G1Jacobian), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaEventRecord(g_msm_time_ev[3], s));
        ZK_CUDA(cudaStreamSynchronize(s));
        if (g_msm_pipe_auto) {
            float ms_copy = 0.f, ms_all = 0.f;
            if (cudaEventElapsedTime(&ms_copy, g_msm_time_ev[0], g_msm_time_ev[1]) == cudaSuccess &&
                cudaEventElapsedTime(&ms_all, g_msm_time_ev[2], g_msm_time_ev[3]) == cudaSuccess && ms_copy > 0.f) {
                const double r = std::min(8.0, std::max(1.0, (double)ms_all / (double)ms_copy));
                g_msm_pipe_ratio = g_msm_pipe_ratio > 0.0 ? 0.5 * (g_msm_pipe_ratio + r) : r;
            } else {
                cudaGetLastError();
            }
        }
        return;
    }
    if (n && !resident)
        ZK_CUDA(cudaMemcpy2DAsync(ds, n * sizeof(Fr), scalars, stride * sizeof(Fr), n * sizeof(Fr), count,
                                  cudaMemcpyHostToDevice, s));
    msm_device(c, ds, n, count, t->d, n, t->table ? &pre : nullptr, dout, s);
    ZK_CUDA(cudaMemcpyAsync(out_xyz, dout, count * sizeof(G1Jacobian), cudaMemcpyDeviceToHost, s));
    ZK_CUDA(cudaStreamSynchronize(s));
}

int b200zk_msm_g1_registered(uint64_t handle, const uint64_t* scalars, size_t n, uint64_t out_xyz[12]) {
    return guarded([&] { msm_registered(handle, scalars, n, 1, n, out_xyz); });
}

int b200zk_msm_g1_registered_many(uint64_t handle, const uint64_t* scalars, size_t stride, size_t count, size_t n,
                                  uint64_t* out_xyz) {
    return guarded([&] { msm_registered(handle, scalars, stride, count, n, out_xyz); });
}

int b200zk_msm_g1_registered_dev(uint64_t handle, const void* d_scalars, size_t stride, size_t count, size_t n,
                                 void* d_out_xyz, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_out_xyz && (n == 0 || d_scalars), "null argument");
        ZK_REQUIRE(count <= 1 || stride >= n, "batch stride smaller than the column");
        ensure_init();
        Context& c = ctx();
        auto it = c.bases.find(handle);
        ZK_REQUIRE(it != c.bases.end(), "unknown bases handle");
        BaseTable* t = it->second;
        ZK_REQUIRE(n <= t->n, "more scalars than registered bases");
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        MsmPre pre{t->table, t->n, t->c, t->nwin};
        msm_device(c, (const Fr*)d_scalars, stride, count, t->d, n, t->table ? &pre : nullptr, (G1Jacobian*)d_out_xyz, s);
    });
}

int b200zk_msm_g1_dev(const void* d_scalars, const void* d_bases, size_t n, uint64_t out_xyz[12], void* stream) {
    return guarded([&] {
        ZK_REQUIRE(out_xyz && (n == 0 || (d_scalars && d_bases)), "null argument");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        G1Jacobian* dout = (G1Jacobian*)c.scratch(s).misc.get(sizeof(G1Jacobian));
        msm_device(c, (const Fr*)d_scalars, n, 1, (const G1Affine*)d_bases, n, nullptr, dout, s);
        copy_point_out(c, dout, out_xyz, s);
    });
}

int b200zk_msm_g1_dev_async(const void* d_scalars, const void* d_bases, size_t n, void* d_out_xyz, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_out_xyz && (n == 0 || (d_scalars && d_bases)), "null argument");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        msm_device(c, (const Fr*)d_scalars, n, 1, (const G1Affine*)d_bases, n, nullptr, (G1Jacobian*)d_out_xyz, s);
    });
}

int b200zk_msm_tune(uint32_t max_chunk, uint32_t max_seglen, uint32_t force_window_bits) {
    return guarded([&] {
        ZK_REQUIRE(max_chunk >= 4 && max_chunk <= 4096, "max_chunk out of range");
        ZK_REQUIRE(max_seglen >= 4 && max_seglen <= 4096, "max_seglen out of range");
        ZK_REQUIRE(force_window_bits == 0 || (force_window_bits >= 4 && force_window_bits <= 23), "window bits out of range");
        g_msm_max_chunk = max_chunk;
        g_msm_max_seglen = max_seglen;
        g_msm_seg_pinned = max_seglen != 64;      // a caller-chosen segment length holds at every size
        g_msm_force_c = force_window_bits;
    });
}

int b200zk_msm_upload_pipeline(uint32_t parts, size_t min_n) {
    return guarded([&] {
        ZK_REQUIRE(parts <= MSM_MAX_PARTS, "parts out of range");
        g_msm_pipe_parts = parts ? parts : MSM_PIPE_PARTS_DEFAULT;      // 0: the library's defaults
        g_msm_pipe_min_n = parts ? min_n : MSM_PIPE_MIN_N_DEFAULT;
        g_msm_pipe_auto = parts == 0 && !getenv("B200ZK_MSM_PIPE_GROWTH") && !getenv("B200ZK_MSM_PIPE_PARTS");
        g_msm_pipe_ratio = 0.0;
    });
}

int b200zk_msm_upload_ranges(size_t n, uint32_t parts, double growth, size_t* begin_out, uint32_t* count_out) {
    return guarded([&] {
        ZK_REQUIRE(begin_out && count_out, "null argument");
        ZK_REQUIRE(parts >= 1 && parts <= MSM_MAX_PARTS, "parts out of range");
        *count_out = (uint32_t)msm_upload_ranges(n, parts, growth, begin_out);
    });
}

int b200zk_msm_profile(int enable) {
    return guarded([&] { g_msm_profile = enable != 0; });
}

int b200zk_msm_last_stages(float* ms_out, int capacity, uint64_t info_out[5]) {
    return guarded([&] {
        ZK_REQUIRE(ms_out && capacity >= MSM_ST_COUNT - 1 && info_out, "bad arguments");
        ZK_REQUIRE(g_msm_last_profiled, "no profiled MSM has run");
        MsmStageEvents* E = g_msm_last_profiled;
        ZK_CUDA(cudaStreamSynchronize(g_msm_ev_stream));
        for (int i = 0; i + 1 < MSM_ST_COUNT; ++i) {
            ms_out[i] = 0.f;
            if (!E->valid[i]) continue;
            int j = i + 1;
            while (j < MSM_ST_COUNT && !E->valid[j]) ++j;
            if (j < MSM_ST_COUNT) ZK_CUDA(cudaEventElapsedTime(&ms_out[i], E->ev[i], E->ev[j]));
        }
        if (E->d_run) {       // the plan the device made for that call (still in the stream's work arena)
            MsmRun h;
            ZK_CUDA(cudaMemcpy(&h, E->d_run, sizeof h, cudaMemcpyDeviceToHost));
            E->info.npairs = h.npairs;
            E->info.chunk = h.L;
        }
        info_out[0] = E->info.n; info_out[1] = E->info.c; info_out[2] = E->info.nwin;
        info_out[3] = E->info.npairs; info_out[4] = E->info.chunk;
    });
}

int b200zk_g1_sum(const uint64_t* points_xyz, size_t count, uint64_t out_xyz[12]) {
    return guarded([&] {
        ZK_REQUIRE(out_xyz && (count == 0 || points_xyz), "null argument");
        ZK_REQUIRE(count < (1u << 20), "too many points");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        G1Jacobian* d = (G1Jacobian*)c.scratch(s).misc.get((count + 1) * sizeof(G1Jacobian));
        if (count) ZK_CUDA(cudaMemcpyAsync(d + 1, points_xyz, count * sizeof(G1Jacobian), cudaMemcpyHostToDevice, s));
        g1_sum_kernel<<<1, 32, 0, s>>>(d + 1, (uint32_t)count, d);
        ZK_LAUNCH_CHECK();
        copy_point_out(c, d, out_xyz, s);
    });
}

}  // extern "C"
