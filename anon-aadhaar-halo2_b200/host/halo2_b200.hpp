// halo2_b200.hpp — C++17 host mirror of the halo2_proofs items whose bodies the b200zk C
// ABI replaces.  The reference's host language is Rust and no Rust toolchain exists in this
// image, so the host side above the C ABI is written in C++ with the same names, argument
// meaning and error behaviour as upstream ([DEP] halo2_proofs 0.2.0 @ v2023_01_20,
// reference Cargo.lock:469-471):
//
//   arithmetic::best_multiexp(coeffs, bases) -> G1          halo2::best_multiexp
//   arithmetic::best_fft(a, omega, log_n)                   halo2::best_fft
//   poly::EvaluationDomain::{new, lagrange_to_coeff,        halo2::EvaluationDomain
//        coeff_to_extended, extended_to_coeff,
//        divide_by_vanishing_poly}
//   poly::kzg::commitment::ParamsKZG::{commit,              halo2::ParamsKZG
//        commit_lagrange}
//
// Upstream's functions are infallible and `assert!` on bad lengths; here a failed
// assertion or a non-zero ABI status throws std::runtime_error (the Rust shim panics).
// Types are the halo2curves wire layouts: Fr = 4 x u64 Montgomery limbs, G1Affine = 8,
// G1 (Jacobian) = 12.  Domain constants are derived on the host exactly as
// EvaluationDomain::new does, using the host bodies of csrc/field.cuh.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200zk.h"
#include "../csrc/field.cuh"

namespace halo2 {

using Fr = std::array<uint64_t, 4>;
using G1Affine = std::array<uint64_t, 8>;
using G1 = std::array<uint64_t, 12>;

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("b200zk: ") + b200zk_last_error());
}
inline void require(bool ok, const char* what) {
    if (!ok) throw std::runtime_error(std::string("assertion failed: ") + what);
}

namespace detail {
inline zk::Fr to_dev(const Fr& a) {
    zk::Fr r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)a[i]; r.l[2 * i + 1] = (uint32_t)(a[i] >> 32); }
    return r;
}
inline Fr from_dev(const zk::Fr& lazy) {
    const zk::Fr a = lazy.canon();   // field.cuh keeps values in [0, 2p); the wire type is reduced
    Fr r;
    for (int i = 0; i < 4; ++i) r[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
    return r;
}
// halo2curves bn256::Fr::ROOT_OF_UNITY (= 7^((r-1)/2^28)) and ZETA, Montgomery form
inline zk::Fr root_of_unity() {
    // canonical 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
    zk::Fr c;
    const uint32_t v[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    for (int i = 0; i < 8; ++i) c.l[i] = v[i];
    return c.to_mont();
}
inline zk::Fr zeta() {
    // canonical 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
    zk::Fr c;
    const uint32_t v[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
    for (int i = 0; i < 8; ++i) c.l[i] = v[i];
    return c.to_mont();
}
inline zk::Fr from_u64(uint64_t x) {
    zk::Fr c = zk::Fr::zero();
    c.l[0] = (uint32_t)x; c.l[1] = (uint32_t)(x >> 32);
    return c.to_mont();
}
}  // namespace detail

// arithmetic.rs best_fft::<Fr>
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    require(a.size() == ((size_t)1 << log_n), "a.len() == 1 << log_n");
    check(b200zk_ntt(a.data()->data(), log_n, omega.data()));
}

// arithmetic.rs best_multiexp::<G1Affine>
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    require(coeffs.size() == bases.size(), "coeffs.len() == bases.len()");
    G1 out{};
    check(b200zk_msm_g1(coeffs.empty() ? nullptr : coeffs.data()->data(), bases.empty() ? nullptr : bases.data()->data(),
                        coeffs.size(), out.data()));
    return out;
}

// poly/domain.rs EvaluationDomain::<Fr>
class EvaluationDomain {
  public:
    uint32_t k, extended_k;
    uint64_t n, quotient_poly_degree;
    Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor;
    std::vector<Fr> t_evaluations;  // inverted, as upstream stores them

    // EvaluationDomain::new(j, k)
    EvaluationDomain(uint32_t j, uint32_t k_) : k(k_) {
        n = (uint64_t)1 << k;
        quotient_poly_degree = j - 1;
        extended_k = k;
        while (((uint64_t)1 << extended_k) < n * quotient_poly_degree) ++extended_k;
        require(extended_k <= 28, "extended_k <= Fr::S");
        zk::Fr w = detail::root_of_unity();
        for (uint32_t i = extended_k; i < 28; ++i) w = w.sqr();
        const zk::Fr ext_w = w;
        for (uint32_t i = k; i < extended_k; ++i) w = w.sqr();
        const zk::Fr z = detail::zeta();
        omega = detail::from_dev(w);
        omega_inv = detail::from_dev(w.inverse());
        extended_omega = detail::from_dev(ext_w);
        extended_omega_inv = detail::from_dev(ext_w.inverse());
        g_coset = detail::from_dev(z);
        g_coset_inv = detail::from_dev(z.sqr());
        ifft_divisor = detail::from_dev(detail::from_u64(n).inverse());
        extended_ifft_divisor = detail::from_dev(detail::from_u64((uint64_t)1 << extended_k).inverse());
        // t(X) = X^n - 1 on the coset: period 2^(extended_k - k)
        const zk::Fr orig = z.pow_u64(n), step = ext_w.pow_u64(n);
        zk::Fr cur = orig;
        do {
            t_evaluations.push_back(detail::from_dev((cur - zk::Fr::one()).inverse()));
            cur = cur * step;
        } while (cur != orig);
        require(t_evaluations.size() == ((size_t)1 << (extended_k - k)), "t_evaluations.len() == 1 << (extended_k - k)");
    }
    size_t extended_len() const { return (size_t)1 << extended_k; }

    std::vector<Fr> lagrange_to_coeff(std::vector<Fr> a) const {
        require(a.size() == n, "a.len() == 1 << k");
        check(b200zk_intt(a.data()->data(), k, omega_inv.data(), ifft_divisor.data()));
        return a;
    }
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        require(a.size() == n, "a.len() == 1 << k");
        std::vector<Fr> out(extended_len());
        check(b200zk_coeff_to_extended(a.data()->data(), k, out.data()->data(), extended_k, extended_omega.data(),
                                       g_coset.data()));
        return out;
    }
    std::vector<Fr> extended_to_coeff(const std::vector<Fr>& a) const {
        require(a.size() == extended_len(), "a.len() == extended_len()");
        std::vector<Fr> out(n * quotient_poly_degree);
        check(b200zk_extended_to_coeff(a.data()->data(), extended_k, extended_omega_inv.data(),
                                       extended_ifft_divisor.data(), g_coset.data(), out.data()->data(), out.size()));
        return out;
    }
    std::vector<Fr> divide_by_vanishing_poly(std::vector<Fr> h) const {
        require(h.size() == extended_len(), "h.len() == extended_len()");
        check(b200zk_divide_by_vanishing(h.data()->data(), extended_k, t_evaluations.data()->data(),
                                         (uint32_t)t_evaluations.size()));
        return h;
    }
};

// poly/kzg/commitment.rs ParamsKZG<Bn256>: the two base tables, uploaded once
class ParamsKZG {
  public:
    ParamsKZG(const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange) : n_(g.size()) {
        require(g.size() == g_lagrange.size(), "g.len() == g_lagrange.len()");
        check(b200zk_bases_register(g.data()->data(), g.size(), &h_g_));
        check(b200zk_bases_register(g_lagrange.data()->data(), g_lagrange.size(), &h_gl_));
    }
    ~ParamsKZG() {
        if (h_g_) b200zk_bases_evict(h_g_);
        if (h_gl_) b200zk_bases_evict(h_gl_);
    }
    ParamsKZG(const ParamsKZG&) = delete;
    ParamsKZG& operator=(const ParamsKZG&) = delete;
    G1 commit(const std::vector<Fr>& poly) const { return msm(h_g_, poly); }
    G1 commit_lagrange(const std::vector<Fr>& poly) const { return msm(h_gl_, poly); }

  private:
    G1 msm(uint64_t h, const std::vector<Fr>& poly) const {
        require(poly.size() <= n_, "bases.len() >= poly.len()");
        G1 out{};
        check(b200zk_msm_g1_registered(h, poly.empty() ? nullptr : poly.data()->data(), poly.size(), out.data()));
        return out;
    }
    size_t n_;
    uint64_t h_g_ = 0, h_gl_ = 0;
};

// ---- the prover loops either side of the hot path (SURVEY.md section 8 f) --------------
namespace detail {
// RAII device column (b200zk_dev_* handle)
struct DevCol {
    uint64_t h = 0;
    explicit DevCol(size_t n) { check(b200zk_dev_alloc(n, &h)); }
    DevCol(const std::vector<Fr>& a) : DevCol(a.size()) {
        if (!a.empty()) check(b200zk_dev_upload(h, 0, a.data()->data(), a.size()));
    }
    ~DevCol() { if (h) b200zk_dev_free(h); }
    DevCol(const DevCol&) = delete;
    DevCol& operator=(const DevCol&) = delete;
    void* ptr() const { return b200zk_dev_ptr(h); }
};
}  // namespace detail

// ff::BatchInvert: `a.iter_mut().batch_invert()`; zeros stay zero
// RAII page-lock of a host vector (b200zk_host_register): the same guard the Rust shim takes around
// the polynomial vectors of a proof, so host-pointer calls transfer at the full PCIe rate and the
// MSM upload / NTT transfer pipelines can overlap their copies.
template <class T> class PageLocked {
public:
    explicit PageLocked(std::vector<T>& v) : p_(v.empty() ? nullptr : v.data()) {
        if (p_) check(b200zk_host_register(p_, v.size() * sizeof(T)));
    }
    ~PageLocked() { if (p_) b200zk_host_unregister(p_); }
    PageLocked(const PageLocked&) = delete;
    PageLocked& operator=(const PageLocked&) = delete;
private:
    void* p_;
};

inline void batch_invert(std::vector<Fr>& a) {
    if (!a.empty()) check(b200zk_batch_invert(a.data()->data(), a.size()));
}

// arithmetic.rs eval_polynomial(poly, point)
inline Fr eval_polynomial(const std::vector<Fr>& poly, const Fr& point) {
    Fr out{};
    if (poly.empty()) return out;
    detail::DevCol d(poly);
    check(b200zk_eval_polynomial_dev(d.ptr(), poly.size(), 1, poly.size(), point.data(), out.data(), nullptr));
    return out;
}

// arithmetic.rs kate_division(a, b): a(X) / (X - b)
inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
    require(!a.empty(), "a.len() >= 1");
    std::vector<Fr> q(a.size() - 1);
    if (q.empty()) return q;
    detail::DevCol d(a), dq(q.size());
    check(b200zk_kate_division_dev(d.ptr(), a.size(), b.data(), dq.ptr(), nullptr));
    check(b200zk_dev_download(dq.h, 0, q.data()->data(), q.size()));
    return q;
}

// halo2curves G1Affine::to_bytes of Jacobian results (e.g. commitments on their way into a transcript)
inline std::vector<std::array<uint8_t, 32>> g1_to_bytes(const std::vector<G1>& points) {
    std::vector<std::array<uint8_t, 32>> out(points.size());
    if (!points.empty()) check(b200zk_g1_to_bytes(points.data()->data(), points.size(), out.data()->data()));
    return out;
}

}  // namespace halo2
