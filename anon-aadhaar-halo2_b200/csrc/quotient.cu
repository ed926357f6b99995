// Quotient numerator h(X) on the extended coset for sm_100a + its C ABI.
//
// Device replacement for halo2_proofs 0.2.0 `plonk::evaluation::Evaluator::evaluate_h`
// and `GraphEvaluator::evaluate` ([DEP] halo2_proofs/src/plonk/evaluation.rs @
// v2023_01_20, reference Cargo.lock:469-471).  Three kernels, one thread per point of
// the extended domain, folding constraints into `values` in exactly upstream's order
// (values = values * y + term), so the result is bit-identical:
//   quotient_graph_kernel        interpreter of the flattened calculation DAG (custom
//                                gates, and the compressed table value of each lookup)
//   quotient_permutation_kernel  the permutation-argument terms
//   quotient_lookup_kernel       the five constraints of one lookup argument
// The only reference pin for this identity is the SquareCircuit verifier,
// reference solidity_verifier_contract/contract.sol:443-505.
// All extended columns are device resident (b200zk_dev_* handles); every column is read
// once per kernel with idx-contiguous (coalesced) accesses, rotations are the same
// accesses shifted by a constant, so the kernels are bound by HBM bytes
// (#columns + 1) * 2^ext_k * 32 B and by one field multiplication per DAG node.
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ntt.cuh"

#include <cstring>
#include <vector>

namespace zk {

enum : uint32_t { SRC_CONSTANT = 0, SRC_INTERMEDIATE, SRC_FIXED, SRC_ADVICE, SRC_INSTANCE, SRC_CHALLENGE,
                  SRC_BETA, SRC_GAMMA, SRC_THETA, SRC_Y, SRC_PREVIOUS };
enum : uint32_t { OP_ADD = 0, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_HORNER, OP_STORE };

constexpr int QMAX_ROT = 32;

struct GraphDev {
    const Fr* constants;
    const int32_t* rotations;
    const b200zk_calc* calcs;
    const b200zk_src* parts;
    uint32_t n_rotations, n_calcs;
    const Fr* const* fixed;
    const Fr* const* advice;
    const Fr* const* instance;
    const Fr* challenges;
    Fr beta, gamma, theta, y;
    uint32_t log_size, log_rot_scale;
    uint32_t begin, count;  // index range evaluated by this launch
    const Fr* previous;  // may be null
    Fr* out;
};

template <int MAXI>
__global__ void __launch_bounds__(128) quotient_graph_kernel(const __grid_constant__ GraphDev G) {
    const uint32_t tix = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t size = 1u << G.log_size;
    if (tix >= G.count) return;
    const uint32_t idx = G.begin + tix;
    Fr inter[MAXI];
    uint32_t rots[QMAX_ROT];
    for (uint32_t r = 0; r < G.n_rotations; ++r) {
        const int32_t off = G.rotations[r] * (int32_t)(1u << G.log_rot_scale);
        rots[r] = (uint32_t)((int32_t)idx + off) & (size - 1u);   // rem_euclid for a power-of-two size
    }
    const Fr prev = G.previous ? ld_fr(G.previous + idx) : Fr::zero();
    auto get = [&](const b200zk_src& s) -> Fr {
        switch (s.kind) {
            case SRC_CONSTANT: return ldg_fr(G.constants + s.a);
            case SRC_INTERMEDIATE: return inter[s.a];
            case SRC_FIXED: return ldg_fr(G.fixed[s.a] + rots[s.b]);
            case SRC_ADVICE: return ldg_fr(G.advice[s.a] + rots[s.b]);
            case SRC_INSTANCE: return ldg_fr(G.instance[s.a] + rots[s.b]);
            case SRC_CHALLENGE: return ldg_fr(G.challenges + s.a);
            case SRC_BETA: return G.beta;
            case SRC_GAMMA: return G.gamma;
            case SRC_THETA: return G.theta;
            case SRC_Y: return G.y;
            default: return prev;
        }
    };
    Fr last = Fr::zero();
    for (uint32_t ci = 0; ci < G.n_calcs; ++ci) {
        const b200zk_calc c = G.calcs[ci];
        Fr v;
        switch (c.op) {
            case OP_ADD: v = get(c.x) + get(c.y); break;
            case OP_SUB: v = get(c.x) - get(c.y); break;
            case OP_MUL: v = get(c.x) * get(c.y); break;
            case OP_SQUARE: { Fr a = get(c.x); v = a * a; break; }
            case OP_DOUBLE: v = get(c.x).dbl(); break;
            case OP_NEGATE: v = get(c.x).neg(); break;
            case OP_HORNER: {
                v = get(c.x);
                const Fr f = get(c.y);
                for (uint32_t p = 0; p < c.parts_len; ++p) v = v * f + get(G.parts[c.parts_off + p]);
                break;
            }
            default: v = get(c.x); break;  // OP_STORE
        }
        inter[c.target] = v;
        last = v;
    }
    st_fr(G.out + idx, last.canon());
}

struct PermDev {
    Fr* values;
    const Fr* const* colvals;   // n_columns
    const Fr* const* sigma;     // n_columns
    const Fr* const* products;  // n_sets
    const Fr* l0;
    const Fr* l_last;
    const Fr* l_active;
    const Fr* tw_lo;            // extended_omega^x tables (shared with the NTT)
    const Fr* tw_hi;
    uint32_t tw_h;
    uint32_t n_columns, n_sets, chunk_len;
    int32_t last_rotation;
    uint32_t log_size, log_rot_scale;
    uint32_t begin, count;
    Fr beta, gamma, y, delta_start, delta;
};

__global__ void __launch_bounds__(128) quotient_permutation_kernel(const __grid_constant__ PermDev P) {
    const uint32_t tix = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t size = 1u << P.log_size;
    if (tix >= P.count) return;
    const uint32_t idx = P.begin + tix;
    const uint32_t rs = 1u << P.log_rot_scale;
    const uint32_t r_next = (idx + rs) & (size - 1u);
    const uint32_t r_last = (uint32_t)((int32_t)idx + P.last_rotation * (int32_t)rs) & (size - 1u);
    const Fr one = Fr::one();
    const Fr l0 = ldg_fr(P.l0 + idx), l_last = ldg_fr(P.l_last + idx), l_active = ldg_fr(P.l_active + idx);
    Fr v = ld_fr(P.values + idx);
    {
        const Fr z0 = ldg_fr(P.products[0] + idx);
        v = v * P.y + (one - z0) * l0;
        const Fr zl = ldg_fr(P.products[P.n_sets - 1] + idx);
        v = v * P.y + (zl * zl - zl) * l_last;
    }
    for (uint32_t s = 1; s < P.n_sets; ++s) {
        const Fr zi = ldg_fr(P.products[s] + idx);
        const Fr zp = ldg_fr(P.products[s - 1] + r_last);
        v = v * P.y + (zi - zp) * l0;
    }
    // beta_term = extended_omega^idx ; current_delta = beta * zeta * beta_term
    Fr w = ldg_fr(P.tw_lo + (idx & ((1u << P.tw_h) - 1u)));
    if (idx >> P.tw_h) w = w * ldg_fr(P.tw_hi + (idx >> P.tw_h));
    Fr current_delta = P.delta_start * w;
    for (uint32_t s = 0; s < P.n_sets; ++s) {
        const uint32_t lo = s * P.chunk_len;
        const uint32_t hi = min(lo + P.chunk_len, P.n_columns);
        Fr left = ldg_fr(P.products[s] + r_next);
        for (uint32_t c = lo; c < hi; ++c) {
            const Fr val = ldg_fr(P.colvals[c] + idx);
            left = left * (val + P.beta * ldg_fr(P.sigma[c] + idx) + P.gamma);
        }
        Fr right = ldg_fr(P.products[s] + idx);
        for (uint32_t c = lo; c < hi; ++c) {
            const Fr val = ldg_fr(P.colvals[c] + idx);
            right = right * (val + current_delta + P.gamma);
            current_delta = current_delta * P.delta;
        }
        v = v * P.y + (left - right) * l_active;
    }
    st_fr(P.values + idx, v.canon());
}

struct LookupDev {
    Fr* values;
    const Fr* table_values;
    const Fr* product;
    const Fr* permuted_input;
    const Fr* permuted_table;
    const Fr* l0;
    const Fr* l_last;
    const Fr* l_active;
    uint32_t log_size, log_rot_scale;
    uint32_t begin, count;
    Fr beta, gamma, y;
};

__global__ void __launch_bounds__(128) quotient_lookup_kernel(const __grid_constant__ LookupDev L) {
    const uint32_t tix = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t size = 1u << L.log_size;
    if (tix >= L.count) return;
    const uint32_t idx = L.begin + tix;
    const uint32_t rs = 1u << L.log_rot_scale;
    const uint32_t r_next = (idx + rs) & (size - 1u);
    const uint32_t r_prev = (idx - rs) & (size - 1u);
    const Fr one = Fr::one();
    const Fr l0 = ldg_fr(L.l0 + idx), l_last = ldg_fr(L.l_last + idx), l_active = ldg_fr(L.l_active + idx);
    const Fr z = ldg_fr(L.product + idx), a = ldg_fr(L.permuted_input + idx), s = ldg_fr(L.permuted_table + idx);
    const Fr a_minus_s = a - s;
    Fr v = ld_fr(L.values + idx);
    v = v * L.y + (one - z) * l0;
    v = v * L.y + (z * z - z) * l_last;
    {
        const Fr zn = ldg_fr(L.product + r_next);
        const Fr t = ldg_fr(L.table_values + idx);
        v = v * L.y + (zn * (a + L.beta) * (s + L.gamma) - z * t) * l_active;
    }
    v = v * L.y + a_minus_s * l0;
    v = v * L.y + a_minus_s * (a - ldg_fr(L.permuted_input + r_prev)) * l_active;
    st_fr(L.values + idx, v.canon());
}

// ------------------------------------------------------------------------ host side
static DevBuffer& buffer_of(Context& c, uint64_t h, size_t min_elems, const char* what) {
    auto it = c.buffers.find(h);
    if (it == c.buffers.end()) throw Error{std::string("b200zk: unknown device handle for ") + what};
    if (it->second.n_elems < min_elems) throw Error{std::string("b200zk: device column too short: ") + what};
    return it->second;
}

static const Fr* col_ptr(Context& c, uint64_t h, size_t size, const char* what) {
    return (const Fr*)buffer_of(c, h, size, what).p;
}

// Upload a pointer table for a list of handles; returns the device array.
struct PtrStager {
    std::vector<const Fr*> host;
    size_t add(Context& c, const uint64_t* handles, uint32_t n, size_t size, const char* what) {
        size_t off = host.size();
        for (uint32_t i = 0; i < n; ++i) host.push_back(col_ptr(c, handles[i], size, what));
        return off;
    }
    // staged through a library-owned page-locked slot (the handle list is the caller's memory)
    const Fr* const* upload(Context& c, cudaStream_t s) {
        Scratch& w = c.scratch(s);
        const Fr** d = (const Fr**)w.quot_ptrs.get((host.size() + 1) * sizeof(const Fr*));
        if (!host.empty()) {
            void* pin = w.staging.begin(host.size() * sizeof(const Fr*));
            memcpy(pin, host.data(), host.size() * sizeof(const Fr*));
            ZK_CUDA(cudaMemcpyAsync(d, pin, host.size() * sizeof(const Fr*), cudaMemcpyHostToDevice, s));
            w.staging.done(s);
        }
        return d;
    }
};

static void check_env(const b200zk_quotient_env* env) {
    ZK_REQUIRE(env, "null env");
    ZK_REQUIRE(env->ext_k >= env->k && env->ext_k <= 28 && env->ext_k >= 1, "bad k / extended_k");
    ZK_REQUIRE(env->n_fixed == 0 || env->fixed, "null fixed handles");
    ZK_REQUIRE(env->n_advice == 0 || env->advice, "null advice handles");
    ZK_REQUIRE(env->n_instance == 0 || env->instance, "null instance handles");
    ZK_REQUIRE(env->n_challenges == 0 || env->challenges, "null challenges");
    const uint64_t size = (uint64_t)1 << env->ext_k;
    ZK_REQUIRE(env->range_begin <= size && env->range_len <= size - env->range_begin, "index range outside the domain");
}

static void env_range(const b200zk_quotient_env* env, uint32_t& begin, uint32_t& count) {
    const uint64_t size = (uint64_t)1 << env->ext_k;
    begin = (uint32_t)env->range_begin;
    count = (uint32_t)(env->range_len ? env->range_len : size - env->range_begin);
}

}  // namespace zk

using namespace zk;

extern "C" {

int b200zk_dev_alloc(size_t n_elems, uint64_t* handle_out) {
    return guarded([&] {
        ZK_REQUIRE(handle_out, "null argument");
        ensure_init();
        Context& c = ctx();
        DevBuffer b;
        b.n_elems = n_elems;
        ZK_CUDA(cudaMalloc(&b.p, std::max<size_t>(n_elems, 1) * sizeof(Fr)));
        const uint64_t h = c.next_handle++;
        c.buffers[h] = b;
        *handle_out = h;
    });
}

int b200zk_dev_free(uint64_t handle) {
    return guarded([&] {
        Context& c = ctx();
        auto it = c.buffers.find(handle);
        ZK_REQUIRE(it != c.buffers.end(), "unknown device handle");
        if (c.ready) {
            cudaSetDevice(c.device);
            cudaStreamSynchronize(c.stream);
        }
        if (it->second.owned) cudaFree(it->second.p);
        c.buffers.erase(it);
    });
}

int b200zk_dev_view(uint64_t parent, size_t offset, size_t n_elems, uint64_t* handle_out) {
    return guarded([&] {
        ZK_REQUIRE(handle_out, "null argument");
        ensure_init();
        Context& c = ctx();
        DevBuffer& b = buffer_of(c, parent, offset + n_elems, "view");
        DevBuffer v;
        v.p = (Fr*)b.p + offset;
        v.n_elems = n_elems;
        v.owned = false;
        const uint64_t h = c.next_handle++;
        c.buffers[h] = v;
        *handle_out = h;
    });
}

int b200zk_dev_upload(uint64_t handle, size_t offset, const uint64_t* host, size_t n_elems) {
    return guarded([&] {
        ensure_init();
        Context& c = ctx();
        DevBuffer& b = buffer_of(c, handle, offset + n_elems, "upload");
        ZK_REQUIRE(host || n_elems == 0, "null argument");
        const void* mirror = n_elems ? c.mirrors.find(host, n_elems) : nullptr;    // null when mirrors are off
        if (mirror)
            ZK_CUDA(cudaMemcpyAsync((Fr*)b.p + offset, mirror, n_elems * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
        else
            ZK_CUDA(cudaMemcpyAsync((Fr*)b.p + offset, host, n_elems * sizeof(Fr), cudaMemcpyHostToDevice, c.stream));
        ZK_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int b200zk_dev_download(uint64_t handle, size_t offset, uint64_t* host, size_t n_elems) {
    return guarded([&] {
        ensure_init();
        Context& c = ctx();
        DevBuffer& b = buffer_of(c, handle, offset + n_elems, "download");
        ZK_REQUIRE(host || n_elems == 0, "null argument");
        if (c.mirrors.enabled && n_elems) {      // the host buffer now equals this device range: keep a mirror
            void* dm = c.mirrors.insert(host, n_elems);
            ZK_CUDA(cudaMemcpyAsync(dm, (Fr*)b.p + offset, n_elems * sizeof(Fr), cudaMemcpyDeviceToDevice, c.stream));
        }
        ZK_CUDA(cudaMemcpyAsync(host, (Fr*)b.p + offset, n_elems * sizeof(Fr), cudaMemcpyDeviceToHost, c.stream));
        ZK_CUDA(cudaStreamSynchronize(c.stream));
    });
}

void* b200zk_dev_ptr(uint64_t handle) {
    Context& c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.buffers.find(handle);
    return it == c.buffers.end() ? nullptr : it->second.p;
}

int b200zk_quotient_graph(const b200zk_graph* g, const b200zk_quotient_env* env, uint64_t previous_handle,
                          uint64_t out_handle) {
    return guarded([&] {
        ZK_REQUIRE(g, "null graph");
        check_env(env);
        ZK_REQUIRE(g->n_rotations <= (uint32_t)QMAX_ROT, "more than 32 distinct rotations");
        ZK_REQUIRE(g->n_intermediates <= 1024, "more than 1024 intermediates");
        ZK_REQUIRE(g->n_calcs == 0 || g->calcs, "null calcs");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const size_t size = (size_t)1 << env->ext_k;
        // validate indices on the host: the kernel trusts them
        auto check_src = [&](const b200zk_src& x) {
            switch (x.kind) {
                case SRC_CONSTANT: ZK_REQUIRE(x.a < g->n_constants, "constant index out of range"); break;
                case SRC_INTERMEDIATE: ZK_REQUIRE(x.a < g->n_intermediates, "intermediate index out of range"); break;
                case SRC_FIXED: ZK_REQUIRE(x.a < env->n_fixed && x.b < g->n_rotations, "fixed query out of range"); break;
                case SRC_ADVICE: ZK_REQUIRE(x.a < env->n_advice && x.b < g->n_rotations, "advice query out of range"); break;
                case SRC_INSTANCE: ZK_REQUIRE(x.a < env->n_instance && x.b < g->n_rotations, "instance query out of range"); break;
                case SRC_CHALLENGE: ZK_REQUIRE(x.a < env->n_challenges, "challenge index out of range"); break;
                default: ZK_REQUIRE(x.kind <= SRC_PREVIOUS, "unknown value source"); break;
            }
        };
        for (uint32_t i = 0; i < g->n_calcs; ++i) {
            const b200zk_calc& cc = g->calcs[i];
            ZK_REQUIRE(cc.op <= OP_STORE, "unknown calculation");
            ZK_REQUIRE(cc.target < g->n_intermediates, "calculation target out of range");
            check_src(cc.x);
            if (cc.op <= OP_MUL || cc.op == OP_HORNER) check_src(cc.y);
            if (cc.op == OP_HORNER) {
                ZK_REQUIRE((uint64_t)cc.parts_off + cc.parts_len <= g->n_parts, "horner parts out of range");
                for (uint32_t p = 0; p < cc.parts_len; ++p) check_src(g->parts[cc.parts_off + p]);
            }
        }
        // stage graph + pointer tables
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
        const size_t o_const = carve((size_t)g->n_constants * sizeof(Fr));
        const size_t o_rot = carve((size_t)g->n_rotations * 4);
        const size_t o_calc = carve((size_t)g->n_calcs * sizeof(b200zk_calc));
        const size_t o_parts = carve((size_t)g->n_parts * sizeof(b200zk_src));
        const size_t o_chal = carve((size_t)env->n_challenges * sizeof(Fr));
        Scratch& w = c.scratch(s);
        char* base = (char*)w.quot_graph.get(off + 256);
        // one page-locked image of the graph, one copy: the caller's arrays (which it may have page-locked
        // itself) are read before the call returns, not by a DMA that is still queued
        char* pin = (char*)w.staging.begin(off + 256);
        auto up = [&](size_t o, const void* src, size_t bytes) {
            if (bytes) memcpy(pin + o, src, bytes);
        };
        up(o_const, g->constants, (size_t)g->n_constants * sizeof(Fr));
        up(o_rot, g->rotations, (size_t)g->n_rotations * 4);
        up(o_calc, g->calcs, (size_t)g->n_calcs * sizeof(b200zk_calc));
        up(o_parts, g->parts, (size_t)g->n_parts * sizeof(b200zk_src));
        up(o_chal, env->challenges, (size_t)env->n_challenges * sizeof(Fr));
        if (off) ZK_CUDA(cudaMemcpyAsync(base, pin, off, cudaMemcpyHostToDevice, s));
        w.staging.done(s);
        PtrStager ps;
        const size_t pf = ps.add(c, env->fixed, env->n_fixed, size, "fixed column");
        const size_t pa = ps.add(c, env->advice, env->n_advice, size, "advice column");
        const size_t pi = ps.add(c, env->instance, env->n_instance, size, "instance column");
        const Fr* const* dptr = ps.upload(c, s);
        GraphDev G;
        G.constants = (const Fr*)(base + o_const);
        G.rotations = (const int32_t*)(base + o_rot);
        G.calcs = (const b200zk_calc*)(base + o_calc);
        G.parts = (const b200zk_src*)(base + o_parts);
        G.n_rotations = g->n_rotations;
        G.n_calcs = g->n_calcs;
        G.fixed = dptr + pf;
        G.advice = dptr + pa;
        G.instance = dptr + pi;
        G.challenges = (const Fr*)(base + o_chal);
        G.beta = fr_from_limbs(env->beta);
        G.gamma = fr_from_limbs(env->gamma);
        G.theta = fr_from_limbs(env->theta);
        G.y = fr_from_limbs(env->y);
        G.log_size = env->ext_k;
        G.log_rot_scale = env->ext_k - env->k;
        G.previous = previous_handle ? col_ptr(c, previous_handle, size, "previous values") : nullptr;
        G.out = (Fr*)buffer_of(c, out_handle, size, "output values").p;
        env_range(env, G.begin, G.count);
        const unsigned blocks = (unsigned)((G.count + 127) / 128);
        if (g->n_intermediates <= 64) quotient_graph_kernel<64><<<blocks, 128, 0, s>>>(G);
        else if (g->n_intermediates <= 256) quotient_graph_kernel<256><<<blocks, 128, 0, s>>>(G);
        else quotient_graph_kernel<1024><<<blocks, 128, 0, s>>>(G);
        ZK_LAUNCH_CHECK();
        // no synchronisation: inputs were staged through library-owned page-locked slots (copied before
        // the call returns), outputs stay on the device, and the next call is ordered behind this one on
        // the library stream (24 lookups = 48 launches per proof: a sync each would cost more than the kernels)
    });
}

int b200zk_quotient_permutation(const b200zk_quotient_env* env, uint64_t values_handle,
                                const uint32_t* column_kind, const uint32_t* column_index,
                                const uint64_t* sigma_handles, uint32_t n_columns,
                                const uint64_t* product_handles, uint32_t n_sets, uint32_t chunk_len,
                                uint32_t blinding_factors, uint64_t l0_handle, uint64_t l_last_handle,
                                uint64_t l_active_row_handle, const uint64_t extended_omega[4],
                                const uint64_t zeta[4], const uint64_t delta[4]) {
    return guarded([&] {
        check_env(env);
        if (n_sets == 0) return;  // `if !sets.is_empty()`
        ZK_REQUIRE(column_kind && column_index && sigma_handles && product_handles, "null argument");
        ZK_REQUIRE(extended_omega && zeta && delta, "null argument");
        ZK_REQUIRE(chunk_len >= 1, "chunk_len must be >= 1");
        ZK_REQUIRE((uint64_t)n_sets * chunk_len >= n_columns && (uint64_t)(n_sets - 1) * chunk_len < std::max(n_columns, 1u),
                   "n_sets does not match ceil(n_columns / chunk_len)");
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const size_t size = (size_t)1 << env->ext_k;
        std::vector<uint64_t> colh(n_columns);
        for (uint32_t j = 0; j < n_columns; ++j) {
            const uint32_t kind = column_kind[j], ix = column_index[j];
            if (kind == SRC_FIXED) { ZK_REQUIRE(ix < env->n_fixed, "permutation fixed column out of range"); colh[j] = env->fixed[ix]; }
            else if (kind == SRC_ADVICE) { ZK_REQUIRE(ix < env->n_advice, "permutation advice column out of range"); colh[j] = env->advice[ix]; }
            else if (kind == SRC_INSTANCE) { ZK_REQUIRE(ix < env->n_instance, "permutation instance column out of range"); colh[j] = env->instance[ix]; }
            else throw Error{"b200zk: permutation column kind must be 2 (fixed), 3 (advice) or 4 (instance)"};
        }
        PtrStager ps;
        const size_t pc = ps.add(c, colh.data(), n_columns, size, "permutation column");
        const size_t psg = ps.add(c, sigma_handles, n_columns, size, "sigma coset");
        const size_t pp = ps.add(c, product_handles, n_sets, size, "permutation product coset");
        const Fr* const* dptr = ps.upload(c, s);
        NttTables* t = ntt_get_tables(c, fr_from_limbs(extended_omega), env->ext_k, s);
        PermDev P;
        P.values = (Fr*)buffer_of(c, values_handle, size, "values").p;
        P.colvals = dptr + pc;
        P.sigma = dptr + psg;
        P.products = dptr + pp;
        P.l0 = col_ptr(c, l0_handle, size, "l0");
        P.l_last = col_ptr(c, l_last_handle, size, "l_last");
        P.l_active = col_ptr(c, l_active_row_handle, size, "l_active_row");
        P.tw_lo = t->tw_lo;
        P.tw_hi = t->tw_hi;
        P.tw_h = t->tw_h;
        P.n_columns = n_columns;
        P.n_sets = n_sets;
        P.chunk_len = chunk_len;
        P.last_rotation = -(int32_t)(blinding_factors + 1);
        P.log_size = env->ext_k;
        P.log_rot_scale = env->ext_k - env->k;
        P.beta = fr_from_limbs(env->beta);
        P.gamma = fr_from_limbs(env->gamma);
        P.y = fr_from_limbs(env->y);
        P.delta_start = P.beta * fr_from_limbs(zeta);
        P.delta = fr_from_limbs(delta);
        env_range(env, P.begin, P.count);
        quotient_permutation_kernel<<<(unsigned)((P.count + 127) / 128), 128, 0, s>>>(P);
        ZK_LAUNCH_CHECK();
    });
}

int b200zk_quotient_lookup(const b200zk_quotient_env* env, uint64_t values_handle, uint64_t table_values_handle,
                           uint64_t product_handle, uint64_t permuted_input_handle,
                           uint64_t permuted_table_handle, uint64_t l0_handle, uint64_t l_last_handle,
                           uint64_t l_active_row_handle) {
    return guarded([&] {
        check_env(env);
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const size_t size = (size_t)1 << env->ext_k;
        LookupDev L;
        L.values = (Fr*)buffer_of(c, values_handle, size, "values").p;
        L.table_values = col_ptr(c, table_values_handle, size, "table values");
        L.product = col_ptr(c, product_handle, size, "lookup product coset");
        L.permuted_input = col_ptr(c, permuted_input_handle, size, "permuted input coset");
        L.permuted_table = col_ptr(c, permuted_table_handle, size, "permuted table coset");
        L.l0 = col_ptr(c, l0_handle, size, "l0");
        L.l_last = col_ptr(c, l_last_handle, size, "l_last");
        L.l_active = col_ptr(c, l_active_row_handle, size, "l_active_row");
        L.log_size = env->ext_k;
        L.log_rot_scale = env->ext_k - env->k;
        L.beta = fr_from_limbs(env->beta);
        L.gamma = fr_from_limbs(env->gamma);
        L.y = fr_from_limbs(env->y);
        env_range(env, L.begin, L.count);
        quotient_lookup_kernel<<<(unsigned)((L.count + 127) / 128), 128, 0, s>>>(L);
        ZK_LAUNCH_CHECK();
    });
}

}  // extern "C"
