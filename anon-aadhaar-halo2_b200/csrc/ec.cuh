// BN254 G1 point arithmetic for the MSM kernels (y^2 = x^3 + 3 over Fq).
//
// Device-side counterpart of halo2curves 0.3.1 `bn256::{G1Affine, G1}`
// ([DEP] halo2curves/src/bn256/curve.rs + src/derive/curve.rs, reference
// Cargo.lock:484-486; curve constant pinned by reference
// solidity_verifier_contract/contract.sol:82).  Wire layouts:
//   G1Affine = {x, y} 64 B Montgomery, identity = (0, 0)
//   G1       = {x, y, z} 96 B Jacobian Montgomery, identity z = 0
//
// Internally buckets are kept in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// identity ZZ = 0): mixed addition costs 8M + 2S with no inversion, and the conversion
// to the Jacobian wire type is (X*ZZ, Y*ZZZ, ZZ), two multiplications.
// All exceptional cases (identity operands, P + P, P + (-P)) are handled exactly so the
// result is the same group element the reference's complete-formula CPU path produces.
#pragma once
#include "field.cuh"

namespace zk {

struct alignas(16) G1Affine {
    Fq x, y;
    ZK_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct alignas(16) G1Jacobian {
    Fq x, y, z;
};

struct alignas(16) G1Xyzz {
    Fq x, y, zz, zzz;

    static ZK_HD G1Xyzz identity() {
        G1Xyzz r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    ZK_HD bool is_identity() const { return zz.is_zero(); }

    static ZK_HD G1Xyzz from_affine(const G1Affine& p) {
        G1Xyzz r;
        if (p.is_identity()) return identity();
        r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }

    // 2 * (affine p), p not the identity
    static ZK_HD G1Xyzz double_affine(const G1Affine& p) {
        G1Xyzz r;
        Fq u = p.y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = p.x * v;
        Fq x2 = p.x.sqr();
        Fq m = x2.dbl() + x2;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * p.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    ZK_HD G1Xyzz dbl() const {
        if (is_identity()) return *this;
        G1Xyzz r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq x2 = x.sqr();
        Fq m = x2.dbl() + x2;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += affine p   (madd-2008-s with the exceptional cases made explicit)
    ZK_HD void add_affine(const G1Affine& p) {
        if (p.is_identity()) return;
        if (is_identity()) {
            x = p.x; y = p.y; zz = Fq::one(); zzz = Fq::one();
            return;
        }
        Fq u2 = p.x * zz;
        Fq s2 = p.y * zzz;
        Fq pp_ = u2 - x;
        Fq r = s2 - y;
        if (pp_.is_zero()) {
            if (r.is_zero()) *this = double_affine(p);
            else *this = identity();
            return;
        }
        Fq pp = pp_.sqr();
        Fq ppp = pp_ * pp;
        Fq q = x * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o   (add-2008-s)
    ZK_HD void add(const G1Xyzz& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        Fq u1 = x * o.zz;
        Fq u2 = o.x * zz;
        Fq s1 = y * o.zzz;
        Fq s2 = o.y * zzz;
        Fq pp_ = u2 - u1;
        Fq r = s2 - s1;
        if (pp_.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        Fq pp = pp_.sqr();
        Fq ppp = pp_ * pp;
        Fq q = u1 * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    ZK_HD G1Xyzz neg() const {
        G1Xyzz r = *this;
        r.y = y.neg();
        return r;
    }

    // halo2curves `G1` wire type; identity -> (0, 1, 0) like G1::identity()
    ZK_HD G1Jacobian to_jacobian() const {
        G1Jacobian j;
        if (is_identity()) {
            j.x = Fq::zero(); j.y = Fq::one(); j.z = Fq::zero();
            return j;
        }
        j.x = (x * zz).canon();
        j.y = (y * zzz).canon();
        j.z = zz.canon();
        return j;
    }

    // affine-normalised Jacobian (z = 1); one field inversion
    ZK_HD G1Jacobian to_jacobian_normalized() const {
        G1Jacobian j;
        if (is_identity()) {
            j.x = Fq::zero(); j.y = Fq::one(); j.z = Fq::zero();
            return j;
        }
        // with zz = z^2, zzz = z^3:  1/zzz = z^-3 ;  1/zz = (z^-3)^2 * z^4 = zi^2 * zz^2
        Fq zi = zzz.inverse();
        Fq zzi = zi.sqr() * zz.sqr();
        j.x = (x * zzi).canon();
        j.y = (y * zi).canon();
        j.z = Fq::one();
        return j;
    }
};

#if defined(__CUDACC__)
// ------------------------------------------------------------------ 4-lane cooperative arithmetic
// The latency-bound tails of the MSM (keyed reduction of a few thousand partial sums, running sums
// over the buckets, the doubling chains of `lo * run`) are chains of dependent point operations; one
// XYZZ addition is 14 field multiplications back to back on one lane (about 8 us on a lone warp).
// A *quad* — four consecutive lanes of a warp holding the same operands — runs the same addition as
// four stages of four independent multiplications, one per lane, with the products exchanged by
// shuffles after every stage: 4 multiplications deep instead of 14.  Every lane of the warp must
// execute these functions together (full-mask shuffles): exceptional operands are handled by
// selecting among results, never by branching around the shuffles.
__device__ __forceinline__ Fq quad_get(const Fq& v, int src_role) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src_role, 4);
    return r;
}
__device__ __forceinline__ Fq fq_sel(bool c, const Fq& a, const Fq& b) {   // c ? a : b
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ Fq fq_sel4(uint32_t role, const Fq& a, const Fq& b, const Fq& c, const Fq& d) {
    return fq_sel(role < 2, fq_sel(role == 0, a, b), fq_sel(role == 2, c, d));
}
__device__ __forceinline__ G1Xyzz xyzz_sel(bool c, const G1Xyzz& a, const G1Xyzz& b) {
    G1Xyzz r;
    r.x = fq_sel(c, a.x, b.x); r.y = fq_sel(c, a.y, b.y); r.zz = fq_sel(c, a.zz, b.zz); r.zzz = fq_sel(c, a.zzz, b.zzz);
    return r;
}

// 2 * a for a quad-replicated a (dbl-2008-s-1, three multiplications deep); identity stays identity.
__device__ __forceinline__ G1Xyzz quad_dbl(const G1Xyzz& a, uint32_t role) {
    const Fq u = a.y.dbl();
    // stage 1: v = u^2 | x2 = x^2
    Fq m1 = fq_sel(role == 0, u, a.x).sqr();
    const Fq v = quad_get(m1, 0), x2 = quad_get(m1, 1);
    const Fq m = x2.dbl() + x2;
    // stage 2: w = u v | s = x v | mm = m^2 | zz3 = v zz
    Fq m2 = fq_sel4(role, u, a.x, m, v) * fq_sel4(role, v, v, m, a.zz);
    const Fq w = quad_get(m2, 0), sv = quad_get(m2, 1), mm = quad_get(m2, 2), zz3 = quad_get(m2, 3);
    const Fq x3 = mm - sv.dbl();
    // stage 3: m (s - x3) | w y | zzz3 = w zzz
    Fq m3 = fq_sel4(role, m, w, w, w) * fq_sel4(role, sv - x3, a.y, a.zzz, a.zzz);
    const Fq t0 = quad_get(m3, 0), t1 = quad_get(m3, 1), zzz3 = quad_get(m3, 2);
    G1Xyzz r;
    r.x = x3; r.y = t0 - t1; r.zz = zz3; r.zzz = zzz3;
    return xyzz_sel(a.is_identity(), a, r);
}

// a + b for quad-replicated operands (add-2008-s, four multiplications deep), all exceptional cases
// exact: identity operands, a = b (doubling) and a = -b (identity).
__device__ __forceinline__ G1Xyzz quad_add(const G1Xyzz& a, const G1Xyzz& b, uint32_t role) {
    // stage 1: u1 = x1 zz2 | u2 = x2 zz1 | s1 = y1 zzz2 | s2 = y2 zzz1
    Fq m1 = fq_sel4(role, a.x, b.x, a.y, b.y) * fq_sel4(role, b.zz, a.zz, b.zzz, a.zzz);
    const Fq u1 = quad_get(m1, 0), u2 = quad_get(m1, 1), s1 = quad_get(m1, 2), s2 = quad_get(m1, 3);
    const Fq P = u2 - u1, R = s2 - s1;
    // stage 2: pp = P^2 | zz12 = zz1 zz2 | rr = R^2 | zzz12 = zzz1 zzz2
    Fq m2 = fq_sel4(role, P, a.zz, R, a.zzz) * fq_sel4(role, P, b.zz, R, b.zzz);
    const Fq pp = quad_get(m2, 0), zz12 = quad_get(m2, 1), rr = quad_get(m2, 2), zzz12 = quad_get(m2, 3);
    // stage 3: ppp = P pp | q = u1 pp | zz3 = zz12 pp
    Fq m3 = fq_sel4(role, P, u1, zz12, zz12) * pp;
    const Fq ppp = quad_get(m3, 0), q = quad_get(m3, 1), zz3 = quad_get(m3, 2);
    const Fq x3 = rr - ppp - q.dbl();
    // stage 4: R (q - x3) | s1 ppp | zzz3 = zzz12 ppp
    Fq m4 = fq_sel4(role, R, s1, zzz12, zzz12) * fq_sel4(role, q - x3, ppp, ppp, ppp);
    const Fq t0 = quad_get(m4, 0), t1 = quad_get(m4, 1), zzz3 = quad_get(m4, 2);
    G1Xyzz r;
    r.x = x3; r.y = t0 - t1; r.zz = zz3; r.zzz = zzz3;
    const bool a_id = a.is_identity(), b_id = b.is_identity();
    const bool same_x = P.is_zero(), same_y = R.is_zero();
    // a = b: the generic formula degenerates; the doubling runs for the whole warp only when some quad needs it
    const bool need_dbl = !a_id && !b_id && same_x && same_y;
    if (__any_sync(0xffffffffu, need_dbl)) {
        const G1Xyzz d = quad_dbl(a, role);
        r = xyzz_sel(need_dbl, d, r);
    }
    r = xyzz_sel(!a_id && !b_id && same_x && !same_y, G1Xyzz::identity(), r);
    r = xyzz_sel(b_id, a, r);
    r = xyzz_sel(a_id && !b_id, b, r);
    return r;
}
#endif  // __CUDACC__

}  // namespace zk
