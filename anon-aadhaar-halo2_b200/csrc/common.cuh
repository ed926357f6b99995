// Library context: device binding, stream, scratch arenas, error reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

// A grow-only device buffer (scratch space that survives across calls).
struct Arena {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) ZK_CUDA(cudaFree(p));
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

struct Context {
    bool ready = false;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;  // the ABI is thread-safe but not concurrent (SURVEY.md section 8 b)
    Arena ntt_io, ntt_tmp, ntt_aux;
    Arena msm_scalars, msm_bases, msm_work, msm_carry;
    Arena misc;
    PinnedArena pinned;
    std::vector<NttTables*> ntt_tables;
    std::map<uint64_t, BaseTable*> bases;
    std::map<uint64_t, DevBuffer> buffers;  // device-resident Fr columns (b200zk_dev_*)
    Arena quot_graph, quot_ptrs;
    Arena poly_work, poly_small, poly_cols, poly_scan;   // csrc/poly.cu scratch
    Arena enc_io;                                          // csrc/encoding.cu staging
    uint64_t next_handle = 1;
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

// Wrap an ABI body: translate exceptions into error codes.
template <class F> int guarded(F&& f) {
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

}  // namespace zk
