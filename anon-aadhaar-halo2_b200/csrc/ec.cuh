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

}  // namespace zk
