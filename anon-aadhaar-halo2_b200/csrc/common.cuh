// Library context: device binding, stream, scratch arenas, error reporting.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "field.cuh"

namespace zk {

struct Error {
    std::string msg;
};

void set_error(const std::string& m);
extern std::atomic<uint64_t> g_launches;

#define ZK_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            char _b[512];                                                                          \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,           \
                     cudaGetErrorString(_e));                                                      \
            throw ::zk::Error{_b};                                                                 \
        }                                                                                          \
    } while (0)

#define ZK_REQUIRE(cond, text)                                                                     \
    do {                                                                                           \
        if (!(cond)) throw ::zk::Error{std::string("b200zk: ") + (text)};                          \
    } while (0)

#define ZK_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        ::zk::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
        ZK_CUDA(cudaGetLastError());                                                               \
    } while (0)

// A grow-only device buffer (scratch space that survives across calls).  Every arena belongs to one
// stream's Scratch, so the only work that can still be using the old block when it has to grow is
// work queued earlier on that stream (or work other streams were ordered behind it): the device is
// drained before the block is released.
struct Arena {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) {
                ZK_CUDA(cudaDeviceSynchronize());
                ZK_CUDA(cudaFree(p));
            }
            p = nullptr; cap = 0;
            size_t want = bytes + bytes / 8;
            ZK_CUDA(cudaMalloc(&p, want));
            cap = want;
        }
        return p;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
};

struct PinnedArena {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) ZK_CUDA(cudaFreeHost(p));
            p = nullptr; cap = 0;
            ZK_CUDA(cudaMallocHost(&p, bytes));
            cap = bytes;
        }
        return p;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
    }
};

struct NttTables;
struct BaseTable;
struct DevBuffer {
    void* p = nullptr;
    size_t n_elems = 0;
    bool owned = true;   // false: a view into another buffer (b200zk_dev_view)
};

// Small host tables (graphs, pointer lists, challenges) travel through page-locked slots owned by
// the library, so a call never leaves a DMA pending on caller memory ("host pointers are never
// retained past the call") and never synchronises for it either: a slot is reused only after the
// copy that read it has completed (its event), 32 uses later.
struct StagingRing {
    static constexpr int SLOTS = 32;
    PinnedArena buf[SLOTS];
    cudaEvent_t ev[SLOTS] = {};
    bool pending[SLOTS] = {};
    int next = 0, cur = -1;
    // a page-locked block of `bytes` for the caller to fill; follow the async copies with done()
    void* begin(size_t bytes) {
        cur = next;
        next = (next + 1) % SLOTS;
        if (pending[cur]) {
            ZK_CUDA(cudaEventSynchronize(ev[cur]));
            pending[cur] = false;
        }
        return buf[cur].get(bytes < 256 ? 256 : bytes);
    }
    void done(cudaStream_t s) {
        if (!ev[cur]) ZK_CUDA(cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming));
        ZK_CUDA(cudaEventRecord(ev[cur], s));
        pending[cur] = true;
    }
    // dst (device) <- src (caller's host memory), asynchronous on s, src read before returning
    void copy(void* dst, const void* src, size_t bytes, cudaStream_t s) {
        if (!bytes) return;
        void* pin = begin(bytes);
        memcpy(pin, src, bytes);
        ZK_CUDA(cudaMemcpyAsync(dst, pin, bytes, cudaMemcpyHostToDevice, s));
        done(s);
    }
    void release() {
        for (int i = 0; i < SLOTS; ++i) {
            if (ev[i]) cudaEventDestroy(ev[i]);
            ev[i] = nullptr;
            pending[i] = false;
            buf[i].release();
        }
    }
};

// Work space of the kernels, one set per CUDA stream the library has been called on: two calls on
// two streams never share a scratch buffer, so the `*_dev` entry points can run concurrently.
struct Scratch {
    Arena ntt_io, ntt_tmp, ntt_aux;
    Arena msm_scalars, msm_bases, msm_work, msm_carry;
    Arena misc;
    Arena quot_graph, quot_ptrs;
    Arena poly_work, poly_small, poly_cols, poly_scan;   // csrc/poly.cu scratch
    Arena enc_io;                                          // csrc/encoding.cu staging
    StagingRing staging;
    void release() {
        for (Arena* a : {&ntt_io, &ntt_tmp, &ntt_aux, &msm_scalars, &msm_bases, &msm_work, &msm_carry, &misc, &quot_graph,
                         &quot_ptrs, &poly_work, &poly_small, &poly_cols, &poly_scan, &enc_io})
            a->release();
        staging.release();
    }
};

// Device mirrors of host Fr buffers (opt-in, b200zk_mirror_enable).  The prover hands the same
// polynomial to several host-pointer calls in a row (commit_lagrange, then lagrange_to_coeff, then
// coeff_to_extended / evaluate_h): with mirrors on, the first call uploads it and the later ones find it
// in HBM; calls that write a host buffer leave its mirror equal to what they wrote.  Coherence is the
// caller's contract: b200zk_mirror_invalidate before it writes or frees a buffer it has passed in (the
// Rust fork does that in `Polynomial`'s `DerefMut` / `Drop`).
struct MirrorCache {
    struct Entry {
        const void* host;
        size_t n_elems;
        void* dev;
        uint64_t last_use;
    };
    bool enabled = false;
    size_t max_bytes = 0, resident_bytes = 0;
    uint64_t clock = 0, hits = 0, misses = 0, evictions = 0;
    std::map<const void*, Entry> entries;                 // keyed by host pointer
    std::multimap<size_t, void*> free_blocks;             // device blocks kept for reuse, by size in bytes

    std::vector<void*> slabs;                             // device blocks come from 64 MB slabs per size class
    void* take_block(size_t bytes) {
        if (!bytes) bytes = 32;
        auto it = free_blocks.find(bytes);
        if (it != free_blocks.end()) {
            void* p = it->second;
            free_blocks.erase(it);
            return p;
        }
        const size_t per = std::max<size_t>(1, ((size_t)64 << 20) / bytes);
        void* slab = nullptr;
        ZK_CUDA(cudaMalloc(&slab, per * bytes));
        slabs.push_back(slab);
        for (size_t i = 1; i < per; ++i) free_blocks.emplace(bytes, (char*)slab + i * bytes);
        return slab;
    }
    void drop(std::map<const void*, Entry>::iterator it) {
        const size_t bytes = it->second.n_elems ? it->second.n_elems * 32 : 32;
        free_blocks.emplace(bytes, it->second.dev);        // stream-ordered reuse on the library stream
        resident_bytes -= it->second.n_elems * 32;
        entries.erase(it);
    }
    // the device copy of host range [host, host + n_elems), or null: an exact mirror, or a slice of a
    // larger one (`&poly[..len]`)
    void* find(const void* host, size_t n_elems) {
        if (!enabled) return nullptr;
        auto it = entries.upper_bound(host);
        if (it != entries.begin()) {
            --it;
            const char* b = (const char*)it->second.host;
            const char* e = b + it->second.n_elems * 32;
            const char* lo = (const char*)host;
            if (lo >= b && lo + n_elems * 32 <= e && ((lo - b) & 31) == 0) {
                it->second.last_use = ++clock;
                ++hits;
                return (char*)it->second.dev + (lo - b);
            }
        }
        ++misses;
        return nullptr;
    }
    // a fresh (uninitialised) mirror for (host, n_elems); the caller fills it with what the host holds
    // (`keep`: a host pointer whose mirror the caller is still using and must survive the eviction)
    void* insert(const void* host, size_t n_elems, const void* keep = nullptr) {
        invalidate(host, n_elems * 32);
        const size_t bytes = n_elems * 32;
        while (resident_bytes + bytes > max_bytes) {
            auto lru = entries.end();
            for (auto it = entries.begin(); it != entries.end(); ++it)
                if (it->first != keep && (lru == entries.end() || it->second.last_use < lru->second.last_use)) lru = it;
            if (lru == entries.end()) break;
            drop(lru);
            ++evictions;
        }
        Entry e{host, n_elems, take_block(bytes), ++clock};
        entries[host] = e;
        resident_bytes += bytes;
        return e.dev;
    }
    // forget every mirror that overlaps [host, host + bytes) (bytes = 0: that contains `host`)
    void invalidate(const void* host, size_t bytes) {
        const char* lo = (const char*)host;
        const char* hi = lo + (bytes ? bytes : 1);
        for (auto it = entries.begin(); it != entries.end();) {
            const char* b = (const char*)it->second.host;
            const char* e = b + it->second.n_elems * 32;
            auto cur = it++;
            if (b < hi && lo < e) drop(cur);
        }
    }
    void release() {
        hits = misses = evictions = 0;
        for (void* p : slabs) cudaFree(p);
        slabs.clear();
        entries.clear();
        free_blocks.clear();
        resident_bytes = 0;
    }
};

struct Context {
    bool ready = false;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;  // the ABI is thread-safe (calls are issued one at a time); the work they queue runs concurrently
    std::map<cudaStream_t, Scratch*> scratch_sets;
    std::vector<NttTables*> ntt_tables;
    std::map<uint64_t, BaseTable*> bases;
    std::map<uint64_t, DevBuffer> buffers;  // device-resident Fr columns (b200zk_dev_*)
    uint64_t next_handle = 1;
    MirrorCache mirrors;
    Scratch& scratch(cudaStream_t s) {
        auto it = scratch_sets.find(s);
        if (it == scratch_sets.end()) it = scratch_sets.emplace(s, new Scratch()).first;
        return *it->second;
    }
    void release_scratch() {
        for (auto& kv : scratch_sets) {
            kv.second->release();
            delete kv.second;
        }
        scratch_sets.clear();
    }
};

Context& ctx();
void ensure_init();
NttTables* ntt_get_tables(Context& c, const Fr& omega, uint32_t log_n, cudaStream_t s);

inline Fr fr_from_limbs(const uint64_t* l) {
    Fr r;
    for (int i = 0; i < 4; ++i) {
        r.l[2 * i] = (uint32_t)l[i];
        r.l[2 * i + 1] = (uint32_t)(l[i] >> 32);
    }
    return r;
}

// Wrap an ABI body: an NVTX range named after the entry point (visible in Nsight Systems / ncu
// --nvtx; costs nothing without a profiler attached), the ABI mutex, exceptions -> error codes.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
template <class F> int guarded_named(const char* name, F&& f) {
    NvtxRange range(name);
    try {
        std::lock_guard<std::mutex> lk(ctx().mu);
        f();
        return 0;
    } catch (const Error& e) {
        set_error(e.msg);
        return 1;
    } catch (const std::exception& e) {
        set_error(std::string("b200zk: ") + e.what());
        return 2;
    }
}
#define guarded(...) guarded_named(__func__, __VA_ARGS__)

}  // namespace zk
