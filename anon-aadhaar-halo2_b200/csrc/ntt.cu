// Host driver + C ABI for the Fr NTT kernels (see ntt.cuh for the algorithm and the
// upstream functions being replaced).
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ntt.cuh"

#include <cstdlib>
#include <cstring>
#include <functional>

namespace zk {

// out[i] = base^(i * mult)   (exponent < 2^32)
static __global__ void fr_pow_table_kernel(Fr base, uint32_t mult, uint32_t count, Fr* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    st_fr(out + i, base.pow_u64((uint64_t)i * mult));
}

static void plan_bits(uint32_t log_n, int& npass, int* bits) {
    npass = (int)((log_n + NTT_MAX_B - 1) / NTT_MAX_B);
    if (npass == 0) npass = 1;
    int base = log_n / npass, rem = log_n % npass;
    for (int p = 0; p < npass; ++p) bits[p] = base + (p < rem ? 1 : 0);
}

static uint32_t g_ntt_direct_tw_max = getenv("B200ZK_NTT_DIRECT_TW_MAX_LOG_N") ? (uint32_t)atoi(getenv("B200ZK_NTT_DIRECT_TW_MAX_LOG_N")) : 20u;

NttTables* ntt_get_tables(Context& c, const Fr& omega, uint32_t log_n, cudaStream_t s) {
    for (NttTables* t : c.ntt_tables)
        if (t->log_n == log_n && t->omega.same_limbs(omega)) return t;
    ZK_REQUIRE(log_n >= 1 && log_n <= 28, "log_n out of range (1..28)");
    NttTables* t = new NttTables();
    t->omega = omega;
    t->log_n = log_n;
    plan_bits(log_n, t->npass, t->bits);
    // inter-pass twiddle omega^e, e = (i_p * J) << log_I: lo / hi tables of sqrt(n) entries (two
    // multiplications), or, where the boundary's 2^(log_n - log_I) distinct exponents fit
    // 2^g_ntt_direct_tw_max entries (L2-resident), one table of exactly those powers (one)
    t->tw_h = (log_n + 1) / 2;
    uint32_t direct_log[4] = {0, 0, 0, 0};
    {
        uint32_t log_I = 0;
        for (int p = 0; p + 1 < t->npass; ++p) {
            if (log_n - log_I <= g_ntt_direct_tw_max) direct_log[p] = log_n - log_I;
            log_I += (uint32_t)t->bits[p];
        }
    }
    const size_t n_lo = (size_t)1 << t->tw_h, n_hi = (size_t)1 << (log_n - t->tw_h);
    size_t total = n_lo + n_hi + 4;
    for (int p = 0; p < 4; ++p)
        if (direct_log[p]) total += (size_t)1 << direct_log[p];
    bool need[NTT_MAX_B + 1] = {};
    for (int p = 0; p < t->npass; ++p) need[t->bits[p]] = true;
    for (int b = 1; b <= NTT_MAX_B; ++b)
        if (need[b]) total += (size_t)1 << b;
    ZK_CUDA(cudaMalloc(&t->block, total * sizeof(Fr)));
    Fr* cur = t->block;
    auto fill = [&](Fr* dst, uint32_t mult, uint32_t count) {
        fr_pow_table_kernel<<<(count + 127) / 128, 128, 0, s>>>(omega, mult, count, dst);
        ZK_LAUNCH_CHECK();
    };
    t->tw_lo = cur; cur += n_lo;
    fill(t->tw_lo, 1, (uint32_t)n_lo);
    t->tw_hi = cur; cur += n_hi;
    fill(t->tw_hi, 1u << t->tw_h, (uint32_t)n_hi);
    for (int b = 1; b <= NTT_MAX_B; ++b) {
        t->tw_tile[b] = nullptr;
        if (!need[b]) continue;
        t->tw_tile[b] = cur; cur += (size_t)1 << b;
        fill(t->tw_tile[b], 1u << (log_n - b), 1u << b);
    }
    {
        uint32_t log_I = 0;
        for (int p = 0; p < 4; ++p) {
            t->tw_direct[p] = nullptr;
            if (p + 1 < t->npass) {
                if (direct_log[p]) {
                    t->tw_direct[p] = cur; cur += (size_t)1 << direct_log[p];
                    fill(t->tw_direct[p], 1u << log_I, 1u << direct_log[p]);
                }
                log_I += (uint32_t)t->bits[p];
            }
        }
    }
    Fr* w8dev = cur;  // 4 entries: omega^(j * n/8), j = 0..3 (n >= 8), else unused
    if (log_n >= 3) {
        fill(w8dev, 1u << (log_n - 3), 4);
        Fr h[4];
        ZK_CUDA(cudaMemcpyAsync(h, w8dev, sizeof h, cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
        t->w8[0] = h[1]; t->w8[1] = h[2]; t->w8[2] = h[3];
    } else {
        // n = 2 or 4: only the size-4 root omega^(n/4) can be needed (as w8[1])
        t->w8[0] = Fr::one();
        t->w8[1] = (log_n == 2) ? omega : Fr::one();
        t->w8[2] = Fr::one();
        ZK_CUDA(cudaStreamSynchronize(s));
    }
    // the tables are complete here (synchronised above), so any stream may use them from now on
    c.ntt_tables.push_back(t);
    return t;
}

// streams / events of the host-buffer transfer pipeline (host_ntt)
constexpr int NTT_PIPE_MAX_CHUNKS = 16;
static cudaStream_t g_ntt_up = nullptr, g_ntt_down = nullptr;
static cudaEvent_t g_ntt_ev_up[NTT_PIPE_MAX_CHUNKS], g_ntt_ev_done[NTT_PIPE_MAX_CHUNKS], g_ntt_ev_start;

static void ntt_drop_tables(Context& c) {
    for (NttTables* t : c.ntt_tables) {
        cudaFree(t->block);
        delete t;
    }
    c.ntt_tables.clear();
}

void ntt_release_tables(Context& c) {
    if (g_ntt_up) {                            // b200zk_shutdown: the next init may bind another device
        cudaStreamDestroy(g_ntt_up);
        cudaStreamDestroy(g_ntt_down);
        for (int i = 0; i < NTT_PIPE_MAX_CHUNKS; ++i) { cudaEventDestroy(g_ntt_ev_up[i]); cudaEventDestroy(g_ntt_ev_done[i]); }
        cudaEventDestroy(g_ntt_ev_start);
        g_ntt_up = g_ntt_down = nullptr;
    }
    ntt_drop_tables(c);
}

// transfer pipeline of the host-buffer entry points (one large transform): column ranges and the
// size from which it is used
static uint32_t g_ntt_pipe_chunks = getenv("B200ZK_NTT_PIPE_CHUNKS") ? (uint32_t)atoi(getenv("B200ZK_NTT_PIPE_CHUNKS")) : 4u;
static uint32_t g_ntt_pipe_min_log_n = getenv("B200ZK_NTT_PIPE_MIN_LOG_N") ? (uint32_t)atoi(getenv("B200ZK_NTT_PIPE_MIN_LOG_N")) : 22u;

struct NttMods {
    uint32_t in_mode = NTT_IN_PLAIN;
    uint32_t n_in = 0;  // 0 = all rows valid
    Fr in_tab[3];
    const Fr* in_table = nullptr;
    uint32_t in_table_mask = 0;
    uint32_t out_mode = NTT_OUT_PLAIN;
    Fr out_tab[3];
    uint64_t n_keep = 0;  // 0 = keep all
};

template <int B, int MODE> static void launch_pass(const NttPassArgs& a, unsigned blocks, unsigned batch, cudaStream_t s) {
    static int configured_device = -1;      // the attribute is per device (b200zk_shutdown + init may rebind)
    constexpr int smem = 2 * NTT_TILE * sizeof(uint4);
    if (configured_device != ctx().device) {
        ZK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<B, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_device = ctx().device;
    }
    ntt_pass_kernel<B, MODE><<<dim3(blocks, batch), NTT_THREADS, smem, s>>>(a);
    ZK_LAUNCH_CHECK();
}

template <int MODE> static void launch_pass_b(int b, const NttPassArgs& a, unsigned blocks, unsigned batch, cudaStream_t s) {
    switch (b) {
        case 1: launch_pass<1, MODE>(a, blocks, batch, s); break;
        case 2: launch_pass<2, MODE>(a, blocks, batch, s); break;
        case 3: launch_pass<3, MODE>(a, blocks, batch, s); break;
        case 4: launch_pass<4, MODE>(a, blocks, batch, s); break;
        case 5: launch_pass<5, MODE>(a, blocks, batch, s); break;
        case 6: launch_pass<6, MODE>(a, blocks, batch, s); break;
        case 7: launch_pass<7, MODE>(a, blocks, batch, s); break;
        case 8: launch_pass<8, MODE>(a, blocks, batch, s); break;
        case 9: launch_pass<9, MODE>(a, blocks, batch, s); break;
        default: throw Error{"b200zk: bad pass width"};
    }
}

// Transform `count` columns of 2^log_n elements.  `in` may equal `out` (in place);
// `tmp` must hold count * 2^log_n elements when more than one pass is needed.
// `chunks` > 1 (single large transform from a host buffer): the first and the last pass are
// launched in `chunks` column ranges; before_first(ch, col0, cols, rows) runs before range ch of
// the first pass is launched, after_last(ch, col0, cols, rows) after range ch of the last pass.
// A range of the first pass reads rows x [col0, col0 + cols) of the row-major [rows][ncols]
// input, a range of the last pass writes the same shape of the output, so the caller can move
// exactly those rectangles over PCIe while the other ranges compute.
struct NttChunkHooks {
    int chunks = 1;
    std::function<void(int, uint64_t, uint64_t, uint64_t)> before_first, after_last;
};

static void ntt_run(Context& c, const Fr* in, uint64_t in_stride, Fr* out, uint64_t out_stride, Fr* tmp,
                    size_t count, uint32_t log_n, const Fr& omega, const NttMods& mods, cudaStream_t s,
                    const NttChunkHooks* hooks = nullptr) {
    if (count == 0) return;
    NttTables* t = ntt_get_tables(c, omega, log_n, s);
    const uint64_t n = (uint64_t)1 << log_n;
    const int P = t->npass;
    // destination of pass p (0-based); see DESIGN.md "NTT buffer rotation"
    auto dst_of = [&](int p) -> Fr* {
        if (P == 1) return out;
        if (P % 2 == 0) return (p % 2 == 0) ? tmp : out;
        if (p == P - 1) return out;      // odd P >= 3: last pass runs in place on `out`
        return (p % 2 == 0) ? tmp : out;
    };
    const Fr* src = in;
    uint64_t src_stride = in_stride;
    uint32_t log_I = 0;
    for (int p = 0; p < P; ++p) {
        const int b = t->bits[p];
        Fr* dst = dst_of(p);
        const uint64_t dst_stride = (dst == tmp) ? n : out_stride;
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        a.in = src;
        a.out = dst;
        a.log_n = log_n;
        a.log_I = log_I;
        a.log_cols = log_n - b;
        a.last = (p == P - 1) ? 1u : 0u;
        a.tw_tile = t->tw_tile[b];
        a.tw_lo = t->tw_lo;
        a.tw_hi = t->tw_hi;
        a.tw_h = t->tw_h;
        a.tw_direct = (p + 1 < P) ? t->tw_direct[p] : nullptr;
        a.w8[0] = t->w8[0]; a.w8[1] = t->w8[1]; a.w8[2] = t->w8[2];
        a.in_mode = (p == 0) ? mods.in_mode : (uint32_t)NTT_IN_PLAIN;
        a.n_in = (p == 0 && mods.n_in) ? mods.n_in : (uint32_t)n;
        for (int i = 0; i < 3; ++i) { a.in_tab[i] = mods.in_tab[i]; a.out_tab[i] = mods.out_tab[i]; }
        a.in_table = mods.in_table;
        a.in_table_mask = mods.in_table_mask;
        a.out_mode = (p == P - 1) ? mods.out_mode : (uint32_t)NTT_OUT_PLAIN;
        a.n_keep = (uint32_t)((p == P - 1 && mods.n_keep) ? mods.n_keep : n);
        a.in_batch_stride = src_stride;
        a.out_batch_stride = dst_stride;
        const uint32_t ncols = 1u << a.log_cols;
        const uint32_t C = 1u << (NTT_TILE_LOG - b);
        const unsigned blocks = (ncols + C - 1) / C;
        const bool chunked = hooks && hooks->chunks > 1 && P >= 2 && (p == 0 || p == P - 1) &&
                             blocks % (unsigned)hooks->chunks == 0;
        if (chunked) {
            const unsigned per = blocks / (unsigned)hooks->chunks;
            for (int ch = 0; ch < hooks->chunks; ++ch) {
                a.block_offset = (uint32_t)ch * per;
                if (p == 0 && hooks->before_first) hooks->before_first(ch, (uint64_t)a.block_offset * C, (uint64_t)per * C, (uint64_t)1 << b);
                if (p == 0) launch_pass_b<1>(b, a, per, (unsigned)count, s);
                else launch_pass_b<0>(b, a, per, (unsigned)count, s);
                if (p == P - 1 && hooks->after_last) hooks->after_last(ch, (uint64_t)a.block_offset * C, (uint64_t)per * C, (uint64_t)1 << b);
            }
        } else if (p == 0) launch_pass_b<1>(b, a, blocks, (unsigned)count, s);
        else launch_pass_b<0>(b, a, blocks, (unsigned)count, s);
        src = dst;
        src_stride = dst_stride;
        log_I += b;
    }
}

static __global__ void fr_scale_table_kernel(Fr* a, size_t n, const Fr* table, uint32_t mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = (ldg_fr(a + i) * ldg_fr(table + (i & mask))).canon();
    st_fr(a + i, x);
}

static void check_batch(size_t stride, size_t count, uint64_t n) {
    ZK_REQUIRE(count <= 65535, "batch count exceeds 65535");
    ZK_REQUIRE(count <= 1 || stride >= n, "batch stride smaller than the transform");
}

// Host-buffer batch: gather columns into one device buffer, transform, scatter back.
static void host_ntt(uint64_t* a, size_t stride, size_t count, uint32_t log_n, const Fr& omega, const NttMods& mods) {
    ensure_init();
    Context& c = ctx();
    const uint64_t n = (uint64_t)1 << log_n;
    check_batch(stride, count, n);
    if (count == 0) return;
    cudaStream_t s = c.stream;
    // a single transform works in the buffer's device mirror when mirrors are on (b200zk_mirror_enable):
    // no upload if an earlier call left the polynomial there, and the result stays for the next one
    Fr* io = nullptr;
    bool resident = false;
    if (count == 1 && c.mirrors.enabled) {
        io = (Fr*)c.mirrors.find(a, n);
        resident = io != nullptr;
        if (!io) io = (Fr*)c.mirrors.insert(a, n);
    } else {
        io = (Fr*)c.scratch(s).ntt_io.get(count * n * sizeof(Fr));
    }
    Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(count * n * sizeof(Fr));
    if (count == 1 && log_n >= g_ntt_pipe_min_log_n && g_ntt_pipe_chunks > 1) {
        // one large transform: upload column ranges under the first pass (unless the polynomial is already
        // resident in its mirror), download under the last
        constexpr int MAXC = NTT_PIPE_MAX_CHUNKS;
        cudaStream_t& up = g_ntt_up;
        cudaStream_t& down = g_ntt_down;
        cudaEvent_t* ev_up = g_ntt_ev_up;
        cudaEvent_t* ev_done = g_ntt_ev_done;
        cudaEvent_t& ev_start = g_ntt_ev_start;
        if (!up) {
            ZK_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
            ZK_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
            for (int i = 0; i < MAXC; ++i) {
                ZK_CUDA(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
                ZK_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
            }
            ZK_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        }
        int chunks = std::min<int>((int)g_ntt_pipe_chunks, MAXC);
        while (chunks & (chunks - 1)) chunks &= chunks - 1;          // power of two
        NttTables* t = ntt_get_tables(c, omega, log_n, s);
        bool uploaded = false;
        if (t->npass >= 2) {
            const uint64_t rows0 = (uint64_t)1 << t->bits[0], ncols0 = n >> t->bits[0];
            // column tiles of the first and of the last pass: both must split into `chunks` ranges
            const int bl = t->bits[t->npass - 1];
            const uint64_t tiles_first = ncols0 >> std::min<uint64_t>(NTT_TILE_LOG - t->bits[0], 63);
            const uint64_t tiles_last = (n >> bl) >> std::min<uint64_t>(NTT_TILE_LOG - bl, 63);
            while (chunks > 1 && ((uint64_t)chunks > tiles_first || (uint64_t)chunks > tiles_last)) chunks >>= 1;
            if (chunks > 1) {
                ZK_CUDA(cudaEventRecord(ev_start, s));              // earlier users of the staging buffers
                ZK_CUDA(cudaStreamWaitEvent(up, ev_start, 0));
                const uint64_t cols = ncols0 / (uint64_t)chunks;
                for (int ch = 0; ch < chunks && !resident; ++ch) {
                    ZK_CUDA(cudaMemcpy2DAsync(io + ch * cols, ncols0 * sizeof(Fr), (const Fr*)a + ch * cols, ncols0 * sizeof(Fr),
                                              cols * sizeof(Fr), rows0, cudaMemcpyHostToDevice, up));
                    ZK_CUDA(cudaEventRecord(ev_up[ch], up));
                }
                uploaded = true;
            }
        }
        if (uploaded) {
            NttChunkHooks hooks;
            hooks.chunks = chunks;
            hooks.before_first = [&](int ch, uint64_t, uint64_t, uint64_t) {
                if (!resident) ZK_CUDA(cudaStreamWaitEvent(s, ev_up[ch], 0));
            };
            hooks.after_last = [&](int ch, uint64_t col0, uint64_t cols, uint64_t rows) {
                const uint64_t ncols = n / rows;
                ZK_CUDA(cudaEventRecord(ev_done[ch], s));
                ZK_CUDA(cudaStreamWaitEvent(down, ev_done[ch], 0));
                ZK_CUDA(cudaMemcpy2DAsync((Fr*)a + col0, ncols * sizeof(Fr), io + col0, ncols * sizeof(Fr), cols * sizeof(Fr), rows,
                                          cudaMemcpyDeviceToHost, down));
            };
            ntt_run(c, io, n, io, n, tmp, 1, log_n, omega, mods, s, &hooks);
            ZK_CUDA(cudaStreamSynchronize(s));
            ZK_CUDA(cudaStreamSynchronize(down));
            return;
        }
    }
    if (count == 1) {           // plain copies: no 2-D descriptor on the prover's per-polynomial path
        if (!resident) ZK_CUDA(cudaMemcpyAsync(io, a, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
        ntt_run(c, io, n, io, n, tmp, 1, log_n, omega, mods, s);
        ZK_CUDA(cudaMemcpyAsync(a, io, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
        return;
    }
    ZK_CUDA(cudaMemcpy2DAsync(io, n * sizeof(Fr), a, stride * sizeof(Fr), n * sizeof(Fr), count,
                              cudaMemcpyHostToDevice, s));
    ntt_run(c, io, n, io, n, tmp, count, log_n, omega, mods, s);
    ZK_CUDA(cudaMemcpy2DAsync(a, stride * sizeof(Fr), io, n * sizeof(Fr), n * sizeof(Fr), count,
                              cudaMemcpyDeviceToHost, s));
    ZK_CUDA(cudaStreamSynchronize(s));
}

// host Fr buffer -> device: its mirror when mirrors are on (uploaded on a miss), else `fallback` + upload
static const Fr* host_in(Context& c, const uint64_t* host, size_t n_elems, Arena& fallback, cudaStream_t s) {
    if (c.mirrors.enabled && n_elems) {
        Fr* d = (Fr*)c.mirrors.find(host, n_elems);
        if (d) return d;
        d = (Fr*)c.mirrors.insert(host, n_elems);
        ZK_CUDA(cudaMemcpyAsync(d, host, n_elems * sizeof(Fr), cudaMemcpyHostToDevice, s));
        return d;
    }
    Fr* d = (Fr*)fallback.get(std::max<size_t>(n_elems, 1) * sizeof(Fr));
    ZK_CUDA(cudaMemcpyAsync(d, host, n_elems * sizeof(Fr), cudaMemcpyHostToDevice, s));
    return d;
}

static NttMods scale_mods(const uint64_t* divisor) {
    NttMods m;
    if (divisor) {
        Fr d = fr_from_limbs(divisor);
        m.out_mode = NTT_OUT_MOD3;
        m.out_tab[0] = d; m.out_tab[1] = d; m.out_tab[2] = d;
    }
    return m;
}

static NttMods coset_in_mods(uint32_t k, const uint64_t* zeta) {
    NttMods m;
    Fr z = fr_from_limbs(zeta);
    m.in_mode = NTT_IN_MOD3;
    m.n_in = 1u << k;
    m.in_tab[0] = Fr::one(); m.in_tab[1] = z; m.in_tab[2] = z * z;  // zeta^(i mod 3)
    return m;
}

static NttMods coset_out_mods(const uint64_t* divisor, const uint64_t* zeta, size_t keep) {
    NttMods m;
    Fr d = fr_from_limbs(divisor);
    Fr z = fr_from_limbs(zeta);
    m.out_mode = NTT_OUT_MOD3;
    // zeta^-(i mod 3): zeta^-1 = zeta^2, zeta^-2 = zeta (zeta^3 = 1)
    m.out_tab[0] = d; m.out_tab[1] = d * (z * z); m.out_tab[2] = d * z;
    m.n_keep = keep;
    return m;
}


// ---------------------------------------------------------------- multi-GPU four-step
// One 2^log_n transform sharded over `world` GPUs (SURVEY.md section 8 e).  With
// n = n1 * n2, input index j = j1 * n2 + j2 and output index i = i1 + n1 * i2:
//   A[i1 + n1 i2] = sum_j2 w^(n1 i2 j2) * w^(i1 j2) * ( sum_j1 a[j1 n2 + j2] * w^(n2 i1 j1) ).
// Rank r owns the columns j2 in [r m, (r+1) m), m = n2 / world, stored [m][n1]; it runs m
// transforms of length n1 (b200zk_ntt_dev), then this kernel multiplies by w^(i1 j2) and
// sends element (i1, j2) to the rank that owns row i1, into its [n1 / world][n2] buffer at
// [i1 mod (n1 / world)][j2].  `dest[s]` is the base of rank s's buffer: either a peer
// mapping (stores travel over NVLink, so the twiddle pass *is* the all-to-all) or a slice
// of a local send buffer for a NCCL all-to-all.  A 32 x 32 tile is transposed through
// shared memory so that both the loads (i1 fastest) and the stores (j2 fastest, 1 KiB
// contiguous per row) are coalesced.
struct Ntt4Args {
    const Fr* in;
    Fr* dest[8];
    uint64_t dest_pitch, dest_col_offset;
    const Fr* tw_lo;
    const Fr* tw_hi;
    uint32_t tw_h;
    uint32_t log_n1, log_m, log_rows;   // rows per destination = n1 / world
    uint32_t col0;                      // r * m
};

static __global__ void __launch_bounds__(256) ntt4_twiddle_scatter_kernel(const __grid_constant__ Ntt4Args A) {
    __shared__ uint4 tile[2][32][33];
    const uint32_t n1 = 1u << A.log_n1, m = 1u << A.log_m;
    const uint32_t i1_0 = blockIdx.x * 32u, jl_0 = blockIdx.y * 32u;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (uint32_t rr = 0; rr < 32; rr += 8) {
        const uint32_t jl = jl_0 + ty + rr, i1 = i1_0 + tx;
        if (jl < m && i1 < n1) {
            Fr x = ldg_fr(A.in + (size_t)jl * n1 + i1);
            const uint32_t e = i1 * (A.col0 + jl);          // < n <= 2^28
            if (e != 0) {
                Fr t = ldg_fr(A.tw_lo + (e & ((1u << A.tw_h) - 1u)));
                const uint32_t eh = e >> A.tw_h;
                if (eh != 0) t = t * ldg_fr(A.tw_hi + eh);
                x = x * t;
            }
            x = x.canon();
            tile[0][ty + rr][tx] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
            tile[1][ty + rr][tx] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t rr = 0; rr < 32; rr += 8) {
        const uint32_t i1 = i1_0 + ty + rr, jl = jl_0 + tx;
        if (jl < m && i1 < n1) {
            const uint32_t s = i1 >> A.log_rows, il = i1 & ((1u << A.log_rows) - 1u);
            uint4* q = reinterpret_cast<uint4*>(A.dest[s] + (size_t)il * A.dest_pitch + A.dest_col_offset + jl);
            q[0] = tile[0][tx][ty + rr];
            q[1] = tile[1][tx][ty + rr];
        }
    }
}


// ------------------------------------------------- cosets of the extended domain (multi-GPU h(X))
// The extended domain {zeta * w_ext^e} splits into P = 2^(ext_k - k) cosets of the base domain:
// e = P j + q  <->  point (zeta w_ext^q) * w^j.  Coset q of a polynomial is an n-point transform of
// its coefficients scaled by (zeta w_ext^q)^i, rotations by w stay inside a coset, so one GPU can
// produce and consume "its" cosets of every column without ever holding an extended column.
static __global__ void fr_scale_pow_kernel(const Fr* __restrict__ in, size_t in_stride, Fr* __restrict__ out,
                                           size_t out_stride, uint32_t n, const Fr* __restrict__ pow_table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (i >= n) return;
    st_fr(out + (size_t)c * out_stride + i, ldg_fr(in + (size_t)c * in_stride + i) * ldg_fr(pow_table + i));
}

// out[c][j] = ext[c][(j << shift) + q]   (gather one coset out of extended columns)
static __global__ void coset_slice_kernel(const Fr* __restrict__ ext, size_t ext_stride, Fr* __restrict__ out,
                                          size_t out_stride, uint32_t n, uint32_t shift, uint32_t q) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (j >= n) return;
    st_fr(out + (size_t)c * out_stride + j, ldg_fr(ext + (size_t)c * ext_stride + (((size_t)j << shift) + q)));
}
// ext[(j << shift) + q] = in[j]   (scatter a coset back into an extended column)
static __global__ void coset_interleave_kernel(const Fr* __restrict__ in, Fr* __restrict__ ext, uint32_t n, uint32_t shift,
                                               uint32_t q) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    st_fr(ext + (((size_t)j << shift) + q), ldg_fr(in + j));
}

}  // namespace zk

using namespace zk;

extern "C" {

int b200zk_ntt_transfer_pipeline(uint32_t chunks, uint32_t min_log_n) {
    return guarded([&] {
        ZK_REQUIRE(chunks >= 1 && chunks <= 16 && (chunks & (chunks - 1)) == 0, "chunks must be a power of two <= 16");
        g_ntt_pipe_chunks = chunks;
        g_ntt_pipe_min_log_n = min_log_n;
    });
}

int b200zk_ntt_tune(uint32_t direct_twiddle_max_log_n) {
    return guarded([&] {
        ZK_REQUIRE(direct_twiddle_max_log_n <= 28, "direct_twiddle_max_log_n out of range (0..28)");
        Context& c = ctx();
        if (c.ready) {                 // cached tables were planned under the old setting
            ZK_CUDA(cudaSetDevice(c.device));
            ZK_CUDA(cudaDeviceSynchronize());
            ntt_drop_tables(c);
        }
        g_ntt_direct_tw_max = direct_twiddle_max_log_n;
    });
}

int b200zk_ntt(uint64_t* a, uint32_t log_n, const uint64_t omega[4]) {
    return b200zk_ntt_many(a, (size_t)1 << log_n, 1, log_n, omega);
}

int b200zk_ntt_many(uint64_t* a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega[4]) {
    return guarded([&] {
        ZK_REQUIRE(a && omega, "null argument");
        if (log_n == 0) return;  // 1-point transform is the identity
        host_ntt(a, stride, count, log_n, fr_from_limbs(omega), NttMods());
    });
}

int b200zk_intt(uint64_t* a, uint32_t log_n, const uint64_t omega_inv[4], const uint64_t divisor[4]) {
    return b200zk_intt_many(a, (size_t)1 << log_n, 1, log_n, omega_inv, divisor);
}

int b200zk_intt_many(uint64_t* a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega_inv[4],
                     const uint64_t divisor[4]) {
    return guarded([&] {
        ZK_REQUIRE(a && omega_inv && divisor, "null argument");
        ZK_REQUIRE(log_n >= 1, "log_n must be >= 1");
        host_ntt(a, stride, count, log_n, fr_from_limbs(omega_inv), scale_mods(divisor));
    });
}

int b200zk_coeff_to_extended(const uint64_t* in, uint32_t k, uint64_t* out, uint32_t ext_k,
                             const uint64_t extended_omega[4], const uint64_t zeta[4]) {
    return b200zk_coeff_to_extended_many(in, (size_t)1 << k, out, (size_t)1 << ext_k, 1, k, ext_k, extended_omega,
                                         zeta);
}

int b200zk_coeff_to_extended_many(const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride,
                                  size_t count, uint32_t k, uint32_t ext_k, const uint64_t extended_omega[4],
                                  const uint64_t zeta[4]) {
    return guarded([&] {
        ZK_REQUIRE(in && out && extended_omega && zeta, "null argument");
        ZK_REQUIRE(ext_k >= k && ext_k >= 1, "extended_k must be >= k and >= 1");
        ensure_init();
        Context& c = ctx();
        const uint64_t n = (uint64_t)1 << k, N = (uint64_t)1 << ext_k;
        check_batch(in_stride, count, n);
        check_batch(out_stride, count, N);
        if (count == 0) return;
        cudaStream_t s = c.stream;
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(count * N * sizeof(Fr));
        const Fr* din = nullptr;
        Fr* io = nullptr;
        if (count == 1 && c.mirrors.enabled) {
            din = host_in(c, in, n, c.scratch(s).ntt_aux, s);
            io = (Fr*)c.mirrors.insert(out, N, in);          // the extended column stays in HBM for evaluate_h
        } else {
            Fr* up = (Fr*)c.scratch(s).ntt_aux.get(count * n * sizeof(Fr));
            io = (Fr*)c.scratch(s).ntt_io.get(count * N * sizeof(Fr));
            ZK_CUDA(cudaMemcpy2DAsync(up, n * sizeof(Fr), in, in_stride * sizeof(Fr), n * sizeof(Fr), count,
                                      cudaMemcpyHostToDevice, s));
            din = up;
        }
        ntt_run(c, din, n, io, N, tmp, count, ext_k, fr_from_limbs(extended_omega), coset_in_mods(k, zeta), s);
        ZK_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(Fr), io, N * sizeof(Fr), N * sizeof(Fr), count,
                                  cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_extended_to_coeff(const uint64_t* a, uint32_t ext_k, const uint64_t extended_omega_inv[4],
                             const uint64_t extended_ifft_divisor[4], const uint64_t zeta[4], uint64_t* out,
                             size_t keep) {
    return guarded([&] {
        ZK_REQUIRE(a && out && extended_omega_inv && extended_ifft_divisor && zeta, "null argument");
        ZK_REQUIRE(ext_k >= 1, "extended_k must be >= 1");
        const uint64_t N = (uint64_t)1 << ext_k;
        ZK_REQUIRE(keep <= N, "keep exceeds the extended domain");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        if (keep == 0) return;
        Fr* io = (Fr*)c.scratch(s).ntt_io.get(N * sizeof(Fr));
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(N * sizeof(Fr));
        const Fr* din = host_in(c, a, N, c.scratch(s).ntt_aux, s);
        NttMods m = coset_out_mods(extended_ifft_divisor, zeta, keep);
        ntt_run(c, din, N, io, N, tmp, 1, ext_k, fr_from_limbs(extended_omega_inv), m, s);
        if (c.mirrors.enabled) {     // the pieces of h(X) are committed next (as slices of `out`): leave them in HBM
            void* dm = c.mirrors.insert(out, keep, a);
            ZK_CUDA(cudaMemcpyAsync(dm, io, keep * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        }
        ZK_CUDA(cudaMemcpyAsync(out, io, keep * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_divide_by_vanishing(uint64_t* h, uint32_t ext_k, const uint64_t* t_evaluations, uint32_t t_len) {
    return guarded([&] {
        ZK_REQUIRE(h && t_evaluations, "null argument");
        ZK_REQUIRE(t_len != 0 && (t_len & (t_len - 1)) == 0, "t_len must be a power of two");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const uint64_t N = (uint64_t)1 << ext_k;
        Fr* io = const_cast<Fr*>(host_in(c, h, N, c.scratch(s).ntt_io, s));   // in place (in the mirror when on)
        Fr* tab = (Fr*)c.scratch(s).ntt_aux.get((size_t)t_len * sizeof(Fr));
        ZK_CUDA(cudaMemcpyAsync(tab, t_evaluations, (size_t)t_len * sizeof(Fr), cudaMemcpyHostToDevice, s));
        fr_scale_table_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(io, N, tab, t_len - 1);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaMemcpyAsync(h, io, N * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_ntt_dev(void* d_a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega[4],
                   const uint64_t* divisor_or_null, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_a && omega, "null argument");
        if (log_n == 0) return;
        ensure_init();
        Context& c = ctx();
        const uint64_t n = (uint64_t)1 << log_n;
        check_batch(stride, count, n);
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(count * n * sizeof(Fr));
        ntt_run(c, (Fr*)d_a, stride, (Fr*)d_a, stride, tmp, count, log_n, fr_from_limbs(omega),
                scale_mods(divisor_or_null), s);
    });
}

int b200zk_coeff_to_extended_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                                 uint32_t k, uint32_t ext_k, const uint64_t extended_omega[4],
                                 const uint64_t zeta[4], void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_in && d_out && extended_omega && zeta, "null argument");
        ZK_REQUIRE(ext_k >= k && ext_k >= 1, "extended_k must be >= k and >= 1");
        ZK_REQUIRE(d_in != d_out, "coeff_to_extended cannot run in place");
        ensure_init();
        Context& c = ctx();
        const uint64_t n = (uint64_t)1 << k, N = (uint64_t)1 << ext_k;
        check_batch(in_stride, count, n);
        check_batch(out_stride, count, N);
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(count * N * sizeof(Fr));
        ntt_run(c, (const Fr*)d_in, in_stride, (Fr*)d_out, out_stride, tmp, count, ext_k,
                fr_from_limbs(extended_omega), coset_in_mods(k, zeta), s);
    });
}

int b200zk_extended_to_coeff_dev(const void* d_a, uint32_t ext_k, const uint64_t extended_omega_inv[4],
                                 const uint64_t extended_ifft_divisor[4], const uint64_t zeta[4],
                                 const void* d_t_evaluations_or_null, uint32_t t_len, void* d_out, size_t keep,
                                 void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_a && d_out && extended_omega_inv && extended_ifft_divisor && zeta, "null argument");
        const uint64_t N = (uint64_t)1 << ext_k;
        ZK_REQUIRE(ext_k >= 1 && keep <= N, "bad extended_k / keep");
        ZK_REQUIRE(d_a != d_out, "extended_to_coeff_dev cannot run in place");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        if (keep == 0) return;
        NttMods m = coset_out_mods(extended_ifft_divisor, zeta, keep);
        if (d_t_evaluations_or_null) {
            ZK_REQUIRE(t_len != 0 && (t_len & (t_len - 1)) == 0, "t_len must be a power of two");
            m.in_mode = NTT_IN_TABLE;
            m.in_table = (const Fr*)d_t_evaluations_or_null;
            m.in_table_mask = t_len - 1;
        }
        // result needs N slots while passes run; d_out may be shorter, so work in scratch
        Fr* io = (Fr*)c.scratch(s).ntt_io.get(N * sizeof(Fr));
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(N * sizeof(Fr));
        ntt_run(c, (const Fr*)d_a, N, io, N, tmp, 1, ext_k, fr_from_limbs(extended_omega_inv), m, s);
        ZK_CUDA(cudaMemcpyAsync(d_out, io, keep * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    });
}

int b200zk_ntt4_twiddle_scatter_dev(const void* d_in, uint32_t log_n, uint32_t log_n1, const uint64_t omega[4],
                                    uint32_t world, uint32_t rank, void* const* dest_bases, size_t dest_pitch,
                                    size_t dest_col_offset, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_in && omega && dest_bases, "null argument");
        ZK_REQUIRE(world >= 1 && world <= 8 && (world & (world - 1)) == 0 && rank < world, "world must be 1, 2, 4 or 8");
        ZK_REQUIRE(log_n >= 2 && log_n <= 28 && log_n1 >= 1 && log_n1 < log_n, "bad transform split");
        uint32_t log_w = 0;
        while ((1u << log_w) < world) ++log_w;
        const uint32_t log_n2 = log_n - log_n1;
        ZK_REQUIRE(log_n1 >= log_w && log_n2 >= log_w, "both factors must be at least the world size");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        NttTables* t = ntt_get_tables(c, fr_from_limbs(omega), log_n, s);
        Ntt4Args a;
        memset(&a, 0, sizeof a);
        a.in = (const Fr*)d_in;
        for (uint32_t i = 0; i < world; ++i) {
            ZK_REQUIRE(dest_bases[i], "null destination buffer");
            a.dest[i] = (Fr*)dest_bases[i];
        }
        a.dest_pitch = dest_pitch;
        a.dest_col_offset = dest_col_offset;
        a.tw_lo = t->tw_lo; a.tw_hi = t->tw_hi; a.tw_h = t->tw_h;
        a.log_n1 = log_n1;
        a.log_m = log_n2 - log_w;
        a.log_rows = log_n1 - log_w;
        a.col0 = rank << a.log_m;
        const uint32_t n1 = 1u << log_n1, m = 1u << a.log_m;
        ntt4_twiddle_scatter_kernel<<<dim3((n1 + 31) / 32, (m + 31) / 32), 256, 0, s>>>(a);
        ZK_LAUNCH_CHECK();
    });
}

int b200zk_ntt4_first_pass_scatter_dev(const void* d_in, uint32_t log_n, uint32_t log_n1, const uint64_t omega[4],
                                       uint32_t world, uint32_t rank, void* const* dest_bases, size_t dest_pitch,
                                       size_t dest_col_offset, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_in && omega && dest_bases, "null argument");
        ZK_REQUIRE(world >= 1 && world <= 8 && (world & (world - 1)) == 0 && rank < world, "world must be 1, 2, 4 or 8");
        ZK_REQUIRE(log_n >= 2 && log_n <= 28 && log_n1 >= 1 && log_n1 < log_n, "bad transform split");
        ZK_REQUIRE(log_n1 <= NTT_MAX_B, "the first factor must fit one pass of the transform kernel");
        uint32_t log_w = 0;
        while ((1u << log_w) < world) ++log_w;
        const uint32_t log_n2 = log_n - log_n1;
        ZK_REQUIRE(log_n1 >= log_w && log_n2 >= log_w, "both factors must be at least the world size");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        // The local step is pass 0 of the whole 2^log_n transform restricted to this rank's columns: n1-point
        // transforms along the rows of the [n1][m] slab (tile twiddles of a 2^log_n1 transform with root omega^n2),
        // the inter-pass twiddle omega^(i1 * j2) of the whole transform, and the store goes to the owner of row i1.
        const Fr w = fr_from_limbs(omega);
        NttTables* big = ntt_get_tables(c, w, log_n, s);
        NttTables* small = ntt_get_tables(c, w.pow_u64((uint64_t)1 << log_n2), log_n1, s);
        ZK_REQUIRE(small->npass == 1 && small->bits[0] == (int)log_n1, "unexpected pass plan");
        const uint32_t log_m = log_n2 - log_w;
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        a.in = (const Fr*)d_in;
        a.out = nullptr;
        a.log_n = log_n;
        a.log_I = 0;
        a.log_cols = log_m;
        a.last = 0;
        a.tw_tile = small->tw_tile[log_n1];
        a.tw_lo = big->tw_lo;
        a.tw_hi = big->tw_hi;
        a.tw_h = big->tw_h;
        a.tw_direct = big->tw_direct[0];
        a.w8[0] = small->w8[0]; a.w8[1] = small->w8[1]; a.w8[2] = small->w8[2];
        a.in_mode = NTT_IN_PLAIN;
        a.n_in = 1u << (log_n1 + log_m);
        a.out_mode = NTT_OUT_PLAIN;
        a.n_keep = a.n_in;
        a.col_base = rank << log_m;
        a.scatter_log_rows = log_n1 - log_w;
        a.scatter_pitch = dest_pitch;
        a.scatter_col_offset = dest_col_offset;
        for (uint32_t i = 0; i < world; ++i) {
            ZK_REQUIRE(dest_bases[i], "null destination buffer");
            a.scatter_dest[i] = (Fr*)dest_bases[i];
        }
        const uint32_t C = 1u << (NTT_TILE_LOG - log_n1);
        const unsigned blocks = ((1u << log_m) + C - 1) / C;
        launch_pass_b<2>((int)log_n1, a, blocks, 1, s);
    });
}

int b200zk_ntt4_gather_rows_dev(const void* d_recv, void* d_out, uint32_t log_n, uint32_t log_n1, uint32_t world,
                                void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_recv && d_out && d_recv != d_out, "bad buffers");
        ZK_REQUIRE(world >= 1 && (world & (world - 1)) == 0, "world must be a power of two");
        uint32_t log_w = 0;
        while ((1u << log_w) < world) ++log_w;
        ZK_REQUIRE(log_n1 < log_n && log_n1 >= log_w && log_n - log_n1 >= log_w, "bad transform split");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        const size_t rows = (size_t)1 << (log_n1 - log_w), n2 = (size_t)1 << (log_n - log_n1), m = n2 / world;
        for (uint32_t r = 0; r < world; ++r)
            ZK_CUDA(cudaMemcpy2DAsync((Fr*)d_out + r * m, n2 * sizeof(Fr), (const Fr*)d_recv + r * rows * m,
                                      m * sizeof(Fr), m * sizeof(Fr), rows, cudaMemcpyDeviceToDevice, s));
    });
}

int b200zk_coeff_to_coset_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                              uint32_t k, const uint64_t omega[4], const uint64_t coset_generator[4], void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_in && d_out && omega && coset_generator, "null argument");
        ZK_REQUIRE(k >= 1 && k <= 28, "k out of range");
        ensure_init();
        Context& c = ctx();
        const uint64_t n = (uint64_t)1 << k;
        check_batch(in_stride, count, n);
        check_batch(out_stride, count, n);
        if (count == 0) return;
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        // powers of the coset generator, one table per call (n entries, ~30 products each)
        Fr* pw = (Fr*)c.scratch(s).ntt_aux.get(n * sizeof(Fr));
        fr_pow_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(fr_from_limbs(coset_generator), 1, (uint32_t)n, pw);
        ZK_LAUNCH_CHECK();
        fr_scale_pow_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)count), 256, 0, s>>>(
            (const Fr*)d_in, in_stride, (Fr*)d_out, out_stride, (uint32_t)n, pw);
        ZK_LAUNCH_CHECK();
        Fr* tmp = (Fr*)c.scratch(s).ntt_tmp.get(count * n * sizeof(Fr));
        ntt_run(c, (Fr*)d_out, out_stride, (Fr*)d_out, out_stride, tmp, count, k, fr_from_limbs(omega), NttMods(), s);
    });
}

int b200zk_extended_coset_slice_dev(const void* d_ext, size_t ext_stride, void* d_out, size_t out_stride, size_t count,
                                    uint32_t k, uint32_t ext_k, uint32_t coset, void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_ext && d_out && d_ext != d_out, "bad buffers");
        ZK_REQUIRE(k >= 1 && ext_k >= k && ext_k <= 28 && coset < (1u << (ext_k - k)), "bad coset");
        ZK_REQUIRE(count <= 65535, "batch count exceeds 65535");
        ensure_init();
        Context& c = ctx();
        if (count == 0) return;
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        const uint32_t n = 1u << k;
        coset_slice_kernel<<<dim3((n + 255) / 256, (unsigned)count), 256, 0, s>>>((const Fr*)d_ext, ext_stride, (Fr*)d_out,
                                                                                   out_stride, n, ext_k - k, coset);
        ZK_LAUNCH_CHECK();
    });
}

int b200zk_extended_coset_interleave_dev(const void* d_coset, void* d_ext, uint32_t k, uint32_t ext_k, uint32_t coset,
                                         void* stream) {
    return guarded([&] {
        ZK_REQUIRE(d_coset && d_ext && d_coset != d_ext, "bad buffers");
        ZK_REQUIRE(k >= 1 && ext_k >= k && ext_k <= 28 && coset < (1u << (ext_k - k)), "bad coset");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        const uint32_t n = 1u << k;
        coset_interleave_kernel<<<(n + 255) / 256, 256, 0, s>>>((const Fr*)d_coset, (Fr*)d_ext, n, ext_k - k, coset);
        ZK_LAUNCH_CHECK();
    });
}

}  // extern "C"
