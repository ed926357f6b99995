// Wire formats of G1 points (SURVEY.md section 8 f, rank 3): what a proof and a params file
// are made of, so that commitments can leave the device already serialised and an SRS can be
// loaded without decompressing 2^k points on the CPU.
//
// Upstream items ([DEP] halo2curves 0.3.1 src/derive/curve.rs `new_curve_impl!`, reference
// Cargo.lock:484-486):
//   G1Affine::to_bytes   32 bytes: x little-endian canonical, bit 7 of byte 31 = parity of the
//                        canonical y; the identity is all-zero.
//   G1Affine::from_bytes the inverse: y = sqrt(x^3 + 3), negated when its parity differs from
//                        the flag; rejects x >= q and non-residues.
// EVM transcript / calldata form (reference solidity_verifier_contract/contract.sol:77-87): 64
// bytes, x then y as 32-byte big-endian canonical integers.
#include "../../include/b200zk.h"
#include "common.cuh"
#include "ec.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace zk {

__device__ __forceinline__ Fq ld_fq_g(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// Jacobian (X, Y, Z) -> affine canonical integers (not Montgomery); identity -> false
__device__ __forceinline__ bool jac_to_canonical_affine(const G1Jacobian& j, Fq& x, Fq& y) {
    if (j.z.is_zero()) return false;
    const Fq zi = j.z.inverse();
    const Fq zi2 = zi.sqr();
    x = (j.x * zi2).from_mont();
    y = (j.y * (zi2 * zi)).from_mont();
    return true;
}

// mode 0: compressed 32 B (to_bytes); mode 1: EVM 64 B big-endian.  is_affine: input stride 64 B
static __global__ void g1_encode_kernel(const Fq* __restrict__ pts, uint32_t count, int is_affine, int mode,
                                        uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fq x, y;
    bool finite;
    if (is_affine) {
        const Fq ax = ld_fq_g(pts + 2 * (size_t)i), ay = ld_fq_g(pts + 2 * (size_t)i + 1);
        finite = !(ax.is_zero() && ay.is_zero());
        x = ax.from_mont();
        y = ay.from_mont();
    } else {
        G1Jacobian j;
        j.x = ld_fq_g(pts + 3 * (size_t)i); j.y = ld_fq_g(pts + 3 * (size_t)i + 1); j.z = ld_fq_g(pts + 3 * (size_t)i + 2);
        finite = jac_to_canonical_affine(j, x, y);
    }
    if (mode == 0) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = finite ? x.l[k] : 0u;
        if (finite) w[7] |= (y.l[0] & 1u) << 31;
        uint4* o = reinterpret_cast<uint4*>(out + 32 * (size_t)i);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {          // big-endian: most significant limb first, bytes swapped
            w[k] = finite ? __byte_perm(x.l[7 - k], 0, 0x0123) : 0u;
            w[8 + k] = finite ? __byte_perm(y.l[7 - k], 0, 0x0123) : 0u;
        }
        uint4* o = reinterpret_cast<uint4*>(out + 64 * (size_t)i);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}

// (q + 1) / 4, the square-root exponent (q = 3 mod 4)
__device__ __forceinline__ Fq fq_sqrt_candidate(const Fq& a) {
    uint32_t e[8];
    Fq::modulus(e);
    // e = (q + 1) >> 2: q + 1 does not overflow 2^256
    uint32_t carry = 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t s = (uint64_t)e[k] + carry;
        e[k] = (uint32_t)s;
        carry = (uint32_t)(s >> 32);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = (e[k] >> 2) | (k < 7 ? e[k + 1] << 30 : 0u);
    return a.pow(e);
}

// status[i]: 0 ok, 1 x not canonical (>= q), 2 not on the curve
static __global__ void g1_decode_kernel(const uint8_t* __restrict__ in, uint32_t count, Fq* __restrict__ out,
                                        uint32_t* __restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint4* p = reinterpret_cast<const uint4*>(in + 32 * (size_t)i);
    const uint4 a = p[0], b = p[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t sign = w[7] >> 31;
    w[7] &= 0x7fffffffu;
    Fq* o = out + 2 * (size_t)i;
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) any |= w[k];
    uint32_t st = 0;
    Fq xm = Fq::zero(), ym = Fq::zero();
    if (any == 0 && sign == 0) {
        // identity
    } else {
        uint32_t m[8], t[8];
        Fq::modulus(m);
        const uint32_t borrow = sub8(t, w, m);
        if (!borrow) st = 1;                      // x >= q
        else {
            Fq x;
#pragma unroll
            for (int k = 0; k < 8; ++k) x.l[k] = w[k];
            xm = x.to_mont();
            Fq three = Fq::one();
            three = three + three + Fq::one();
            const Fq rhs = xm.sqr() * xm + three;
            Fq y = fq_sqrt_candidate(rhs);
            if (y.sqr() != rhs) st = 2;
            else {
                const Fq yc = y.from_mont();
                if ((yc.l[0] & 1u) != sign) y = y.neg();
                ym = y.canon();
                xm = xm.canon();
            }
        }
    }
    if (st != 0) { xm = Fq::zero(); ym = Fq::zero(); }
    uint4* q = reinterpret_cast<uint4*>(o);
    q[0] = make_uint4(xm.l[0], xm.l[1], xm.l[2], xm.l[3]);
    q[1] = make_uint4(xm.l[4], xm.l[5], xm.l[6], xm.l[7]);
    q[2] = make_uint4(ym.l[0], ym.l[1], ym.l[2], ym.l[3]);
    q[3] = make_uint4(ym.l[4], ym.l[5], ym.l[6], ym.l[7]);
    status[i] = st;
}

static void encode(const uint64_t* pts, size_t count, int is_affine, int mode, uint8_t* out) {
    ZK_REQUIRE((pts && out) || count == 0, "null argument");
    ZK_REQUIRE(count < ((size_t)1 << 31), "too many points");
    if (count == 0) return;
    ensure_init();
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const size_t in_bytes = count * (is_affine ? 64 : 96), out_bytes = count * (mode ? 64 : 32);
    char* d = (char*)c.scratch(s).enc_io.get(in_bytes + out_bytes + 256);
    char* dout = d + (in_bytes + 255) / 256 * 256;
    ZK_CUDA(cudaMemcpyAsync(d, pts, in_bytes, cudaMemcpyHostToDevice, s));
    g1_encode_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>((const Fq*)d, (uint32_t)count, is_affine, mode,
                                                                     (uint8_t*)dout);
    ZK_LAUNCH_CHECK();
    ZK_CUDA(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, s));
    ZK_CUDA(cudaStreamSynchronize(s));
}

}  // namespace zk

using namespace zk;

extern "C" {

int b200zk_g1_to_bytes(const uint64_t* points_xyz, size_t count, uint8_t* out32) {
    return guarded([&] { encode(points_xyz, count, 0, 0, out32); });
}

int b200zk_g1_affine_to_bytes(const uint64_t* points_xy, size_t count, uint8_t* out32) {
    return guarded([&] { encode(points_xy, count, 1, 0, out32); });
}

int b200zk_g1_to_evm_bytes(const uint64_t* points_xyz, size_t count, uint8_t* out64) {
    return guarded([&] { encode(points_xyz, count, 0, 1, out64); });
}

int b200zk_g1_affine_from_bytes(const uint8_t* in32, size_t count, uint64_t* points_xy) {
    return guarded([&] {
        ZK_REQUIRE((in32 && points_xy) || count == 0, "null argument");
        ZK_REQUIRE(count < ((size_t)1 << 31), "too many points");
        if (count == 0) return;
        ensure_init();
        Context& c = ctx();
        cudaStream_t s = c.stream;
        const size_t in_bytes = (count * 32 + 255) / 256 * 256, out_bytes = (count * 64 + 255) / 256 * 256;
        char* d = (char*)c.scratch(s).enc_io.get(in_bytes + out_bytes + count * 4);
        uint8_t* din = (uint8_t*)d;
        Fq* dout = (Fq*)(d + in_bytes);
        uint32_t* dst = (uint32_t*)(d + in_bytes + out_bytes);
        ZK_CUDA(cudaMemcpyAsync(din, in32, count * 32, cudaMemcpyHostToDevice, s));
        g1_decode_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(din, (uint32_t)count, dout, dst);
        ZK_LAUNCH_CHECK();
        std::vector<uint32_t> st(count);
        ZK_CUDA(cudaMemcpyAsync(points_xy, dout, count * 64, cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaMemcpyAsync(st.data(), dst, count * 4, cudaMemcpyDeviceToHost, s));
        ZK_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < count; ++i) {
            if (st[i] == 1) throw Error{"b200zk: point " + std::to_string(i) + ": x coordinate is not canonical"};
            if (st[i] == 2) throw Error{"b200zk: point " + std::to_string(i) + ": not on the curve"};
        }
    });
}

}  // extern "C"
