// Context lifetime, error reporting, synthetic-input generators and the modmul-peak
// microbenchmark of the b200zk C ABI (include/b200zk.h).
#include "../../include/b200zk.h"
#include "common.cuh"
#include <cstdlib>
#include <set>
#include "ec.cuh"
#include "ntt.cuh"

namespace zk {

static thread_local std::string t_error;
static std::string g_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& m) {
    t_error = m;
    g_error = m;
}

Context& ctx() {
    static Context c;
    return c;
}

static void init_locked(int device) {
    Context& c = ctx();
    if (c.ready) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        throw Error{std::string("b200zk: no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback"};
    }
    if (device < 0) ZK_CUDA(cudaGetDevice(&device));
    ZK_REQUIRE(device < count, "device index out of range");
    ZK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, device));
    ZK_REQUIRE(prop.major >= 10, "b200zk kernels are built for sm_100a only");
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    ZK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ready = true;
}

void ensure_init() {
    if (!ctx().ready) init_locked(-1);
    else ZK_CUDA(cudaSetDevice(ctx().device));
}

void ntt_release_tables(Context& c);
void msm_release_bases(Context& c);

// ------------------------------------------------------------------ input generators
// Streams defined by oracle/bn254.py (seeded_fr_mont_limbs / seeded_g1_points).
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void gen_scalars_kernel(Fr* out, size_t n, uint64_t seed, size_t start) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint64_t w = splitmix64((seed << 32) + 4 * (start + i) + j);
        if (j == 3) w &= (1ull << 62) - 1;
        v.l[2 * j] = (uint32_t)w;
        v.l[2 * j + 1] = (uint32_t)(w >> 32);
    }
    v.reduce_once();  // value < 2^254 < 2r
    st_fr(out + i, v);
}

// P_i = [t_i] G, t_i = splitmix64(seed << 32 + i) | 1, by a fixed-base table of 2^j G.
__global__ void gen_pow2_table_kernel(G1Affine* tbl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Xyzz p;
    p.x = Fq::one();
    p.y = Fq::one().dbl();
    p.zz = Fq::one();
    p.zzz = Fq::one();
    for (int j = 0; j < 64; ++j) {
        G1Jacobian a = p.to_jacobian_normalized();
        tbl[j].x = a.x;
        tbl[j].y = a.y;
        p = p.dbl();
    }
}

__global__ void gen_points_kernel(G1Affine* out, size_t n, uint64_t seed, size_t start, const G1Affine* tbl) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t t = splitmix64((seed << 32) + (start + i)) | 1ull;
    G1Xyzz acc = G1Xyzz::identity();
    for (int j = 0; j < 64; ++j) {
        if ((t >> j) & 1ull) acc.add_affine(tbl[j]);
    }
    G1Jacobian a = acc.to_jacobian_normalized();
    G1Affine r;
    r.x = a.x;
    r.y = a.y;
    out[i] = r;
}

template <class F>
static __global__ void __launch_bounds__(128) field_op_kernel(uint32_t op, const F* __restrict__ a, const F* __restrict__ b,
                                                              size_t n, F* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F x = a[i];
    F r = F::zero();
    switch (op) {
        case B200ZK_FIELD_MUL: r = x * b[i]; break;
        case B200ZK_FIELD_ADD: r = x + b[i]; break;
        case B200ZK_FIELD_SUB: r = x - b[i]; break;
        case B200ZK_FIELD_SQR: r = x.sqr(); break;
        case B200ZK_FIELD_INV: r = x.inverse(); break;
        case B200ZK_FIELD_NEG: r = x.neg(); break;
        case B200ZK_FIELD_DBL: r = x.dbl(); break;
        case B200ZK_FIELD_FROM_MONT: r = x.from_mont(); break;
        case B200ZK_FIELD_TO_MONT: r = x.to_mont(); break;
        case B200ZK_FIELD_POW: r = x.pow(b[i].l); break;        // exponent: the 256 bits of b[i] as an integer
        default: break;
    }
    out[i] = r.canon();
}


// ---------------------------------------------------------------- modmul peak probe
__global__ void __launch_bounds__(256) modmul_peak_kernel(Fq* out, uint32_t iters) {
    // four independent dependency chains per thread
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a = Fq::one(), b = Fq::r2(), c = Fq::one().dbl(), d = Fq::r2().dbl();
    a.l[0] ^= tid & 0xffu;
    Fq m = Fq::r2();
    m.l[1] ^= tid & 0xfu;
    m.reduce_once();
    a.reduce_once();
    for (uint32_t i = 0; i < iters; ++i) {
        a = a * m;
        b = b * m;
        c = c * m;
        d = d * m;
    }
    Fq s = (a + b) + (c + d);
    if (s.l[0] == 0xdeadbeefu && s.l[7] == 0x12345678u) out[tid] = s;  // keep the chains alive
}

}  // namespace zk

using namespace zk;

// page-locked host memory handed out / registered through the ABI (b200zk_host_*)
static std::set<void*> g_host_registered, g_host_allocated;

extern "C" {

uint32_t b200zk_abi_version(void) { return 1; }

const char* b200zk_last_error(void) { return g_error.c_str(); }

int b200zk_init(int device) {
    return guarded([&] {
        Context& c = ctx();
        if (c.ready) {
            ZK_REQUIRE(device < 0 || device == c.device, "already initialised on another device");
            return;
        }
        init_locked(device);
    });
}

int b200zk_shutdown(void) {
    return guarded([&] {
        Context& c = ctx();
        if (!c.ready) return;
        cudaSetDevice(c.device);
        cudaStreamSynchronize(c.stream);
        ntt_release_tables(c);
        msm_release_bases(c);
        c.release_scratch();
        c.mirrors.release();
        c.mirrors.enabled = false;
        for (auto& kv : c.buffers)
            if (kv.second.owned) cudaFree(kv.second.p);
        c.buffers.clear();
        for (void* p : g_host_registered) cudaHostUnregister(p);
        for (void* p : g_host_allocated) cudaFreeHost(p);
        g_host_registered.clear();
        g_host_allocated.clear();
        cudaStreamDestroy(c.stream);
        c.stream = nullptr;
        c.ready = false;
    });
}

uint64_t b200zk_kernel_launches(void) { return g_launches.load(); }

int b200zk_mirror_enable(size_t max_bytes) {
    return guarded([&] {
        Context& c = ctx();
        if (max_bytes == 0) {
            if (c.ready) {
                ZK_CUDA(cudaSetDevice(c.device));
                ZK_CUDA(cudaDeviceSynchronize());
            }
            c.mirrors.release();
            c.mirrors.enabled = false;
            c.mirrors.max_bytes = 0;
            return;
        }
        ensure_init();
        c.mirrors.enabled = true;
        c.mirrors.max_bytes = max_bytes;
    });
}

int b200zk_mirror_invalidate(const void* host_ptr, size_t bytes) {
    return guarded([&] {
        Context& c = ctx();
        if (!c.mirrors.enabled) return;
        if (!host_ptr) {                           // forget every mirror; the device blocks stay pooled
            while (!c.mirrors.entries.empty()) c.mirrors.drop(c.mirrors.entries.begin());
            c.mirrors.hits = c.mirrors.misses = c.mirrors.evictions = 0;
            return;
        }
        // the library stream may still be reading the mirror (host-pointer calls are synchronous, so it is
        // idle here); freed blocks are only ever reused by work queued later on that stream
        c.mirrors.invalidate(host_ptr, bytes);
    });
}

int b200zk_mirror_stats(uint64_t out[4]) {
    return guarded([&] {
        ZK_REQUIRE(out, "null argument");
        Context& c = ctx();
        out[0] = c.mirrors.hits; out[1] = c.mirrors.misses; out[2] = c.mirrors.resident_bytes; out[3] = c.mirrors.evictions;
    });
}

int b200zk_stream_release(void* stream) {
    return guarded([&] {
        Context& c = ctx();
        if (!c.ready) return;
        cudaStream_t s = stream ? (cudaStream_t)stream : c.stream;
        auto it = c.scratch_sets.find(s);
        if (it == c.scratch_sets.end()) return;
        ZK_CUDA(cudaSetDevice(c.device));
        ZK_CUDA(cudaDeviceSynchronize());
        it->second->release();
        delete it->second;
        c.scratch_sets.erase(it);
    });
}

int b200zk_host_register(void* ptr, size_t bytes) {
    return guarded([&] {
        ZK_REQUIRE(ptr && bytes, "null or empty host buffer");
        ensure_init();
        ZK_REQUIRE(!g_host_registered.count(ptr) && !g_host_allocated.count(ptr), "host buffer already page-locked");
        ZK_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
        g_host_registered.insert(ptr);
    });
}

int b200zk_host_unregister(void* ptr) {
    return guarded([&] {
        ZK_REQUIRE(g_host_registered.count(ptr), "host buffer was not registered");
        cudaStreamSynchronize(ctx().stream);
        ZK_CUDA(cudaHostUnregister(ptr));
        g_host_registered.erase(ptr);
    });
}

int b200zk_host_alloc(size_t bytes, void** ptr_out) {
    return guarded([&] {
        ZK_REQUIRE(ptr_out && bytes, "null argument or empty allocation");
        ensure_init();
        void* p = nullptr;
        ZK_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
        g_host_allocated.insert(p);
        *ptr_out = p;
    });
}

int b200zk_host_free(void* ptr) {
    return guarded([&] {
        ZK_REQUIRE(g_host_allocated.count(ptr), "not a b200zk_host_alloc allocation");
        cudaStreamSynchronize(ctx().stream);
        ZK_CUDA(cudaFreeHost(ptr));
        g_host_allocated.erase(ptr);
    });
}

int b200zk_gen_scalars_dev(void* d_out, size_t n, uint64_t seed, size_t start) {
    return guarded([&] {
        ensure_init();
        if (n == 0) return;
        Context& c = ctx();
        gen_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>((Fr*)d_out, n, seed, start);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int b200zk_gen_points_dev(void* d_out, size_t n, uint64_t seed, size_t start) {
    return guarded([&] {
        ensure_init();
        if (n == 0) return;
        Context& c = ctx();
        G1Affine* tbl = (G1Affine*)c.scratch(c.stream).misc.get(64 * sizeof(G1Affine));
        gen_pow2_table_kernel<<<1, 32, 0, c.stream>>>(tbl);
        ZK_LAUNCH_CHECK();
        gen_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c.stream>>>((G1Affine*)d_out, n, seed, start, tbl);
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaStreamSynchronize(c.stream));
    });
}

// Element-wise field arithmetic on the device (the known-answer test of SURVEY.md section 7 step 3): out[i] =
// a[i] op b[i], operands taken as given (any representative in [0, 2p), as values are between kernels), the
// result canonical.  A diagnostic entry point: the product kernels inline the same Fp<P> operators.
int b200zk_field_op(uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
    return guarded([&] {
        ZK_REQUIRE(field <= 1 && op <= B200ZK_FIELD_OP_LAST, "unknown field or operation");
        ZK_REQUIRE(n == 0 || (a && out), "null argument");
        const bool binary = op == B200ZK_FIELD_MUL || op == B200ZK_FIELD_ADD || op == B200ZK_FIELD_SUB || op == B200ZK_FIELD_POW;
        ZK_REQUIRE(n == 0 || !binary || b, "null second operand");
        if (n == 0) return;
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const size_t bytes = n * 32;
        char* d = (char*)c.scratch(s).misc.get(3 * bytes);
        ZK_CUDA(cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, s));
        if (binary) ZK_CUDA(cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyHostToDevice, s));
        const unsigned blocks = (unsigned)((n + 127) / 128);
        if (field == 0) field_op_kernel<Fr><<<blocks, 128, 0, s>>>(op, (const Fr*)d, (const Fr*)(d + bytes), n, (Fr*)(d + 2 * bytes));
        else field_op_kernel<Fq><<<blocks, 128, 0, s>>>(op, (const Fq*)d, (const Fq*)(d + bytes), n, (Fq*)(d + 2 * bytes));
        ZK_LAUNCH_CHECK();
        ZK_CUDA(cudaMemcpyAsync(out, d + 2 * bytes, bytes, cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200zk_modmul_peak(uint32_t iters, double* modmul_per_s_out) {
    return guarded([&] {
        ensure_init();
        Context& c = ctx();
        const int blocks = c.sm_count * 8, threads = 256;
        Fq* sink = (Fq*)c.scratch(c.stream).misc.get((size_t)blocks * threads * sizeof(Fq));
        cudaEvent_t e0, e1;
        ZK_CUDA(cudaEventCreate(&e0));
        ZK_CUDA(cudaEventCreate(&e1));
        modmul_peak_kernel<<<blocks, threads, 0, c.stream>>>(sink, iters / 8 + 1);  // warm-up
        ZK_LAUNCH_CHECK();
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            ZK_CUDA(cudaEventRecord(e0, c.stream));
            modmul_peak_kernel<<<blocks, threads, 0, c.stream>>>(sink, iters);
            ZK_LAUNCH_CHECK();
            ZK_CUDA(cudaEventRecord(e1, c.stream));
            ZK_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *modmul_per_s_out = 4.0 * iters * (double)blocks * threads / (best * 1e-3);
    });
}

}  // extern "C"
