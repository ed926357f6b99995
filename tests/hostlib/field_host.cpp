// Host build of csrc/field.cuh (its plain-C emulation bodies) for CPU unit tests.
#include "field.cuh"
#include <cstring>
using namespace zk;
template <class F> static F ld(const uint32_t* p) { F f; memcpy(f.l, p, 32); return f; }
// field.cuh keeps values in [0, 2p): `st` stores the canonical representative, `st_raw` the
// lazy one (the tests check raw < 2p and raw == canonical mod p).
template <class F> static void st(uint32_t* p, const F& f) { F c = f.canon(); memcpy(p, c.l, 32); }
template <class F> static void st_raw(uint32_t* p, const F& f) { memcpy(p, f.l, 32); }
extern "C" {
void h_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fr>(a) * ld<Fr>(b)); }
void h_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fr>(a) + ld<Fr>(b)); }
void h_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fr>(a) - ld<Fr>(b)); }
void h_fr_neg(const uint32_t* a, uint32_t* r) { st(r, ld<Fr>(a).neg()); }
void h_fr_inv(const uint32_t* a, uint32_t* r) { st(r, ld<Fr>(a).inverse()); }
void h_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fq>(a) * ld<Fq>(b)); }
void h_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fq>(a) + ld<Fq>(b)); }
void h_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, ld<Fq>(a) - ld<Fq>(b)); }
void h_fq_neg(const uint32_t* a, uint32_t* r) { st(r, ld<Fq>(a).neg()); }
void h_fq_inv(const uint32_t* a, uint32_t* r) { st(r, ld<Fq>(a).inverse()); }
void h_fr_sqr_raw(const uint32_t* a, uint32_t* r) { st_raw(r, ld<Fr>(a).sqr()); }
void h_fq_sqr_raw(const uint32_t* a, uint32_t* r) { st_raw(r, ld<Fq>(a).sqr()); }
void h_fr_mul_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fr>(a) * ld<Fr>(b)); }
void h_fq_mul_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fq>(a) * ld<Fq>(b)); }
void h_fr_add_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fr>(a) + ld<Fr>(b)); }
void h_fq_add_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fq>(a) + ld<Fq>(b)); }
void h_fr_sub_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fr>(a) - ld<Fr>(b)); }
void h_fq_sub_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { st_raw(r, ld<Fq>(a) - ld<Fq>(b)); }
int h_fr_is_zero(const uint32_t* a) { return ld<Fr>(a).is_zero() ? 1 : 0; }
int h_fq_is_zero(const uint32_t* a) { return ld<Fq>(a).is_zero() ? 1 : 0; }
int h_fr_eq(const uint32_t* a, const uint32_t* b) { return ld<Fr>(a) == ld<Fr>(b) ? 1 : 0; }
void h_fr_one(uint32_t* r) { st(r, Fr::one()); }
void h_fq_one(uint32_t* r) { st(r, Fq::one()); }
void h_fr_to_mont(const uint32_t* a, uint32_t* r) { st(r, ld<Fr>(a).to_mont()); }
void h_fr_from_mont(const uint32_t* a, uint32_t* r) { st(r, ld<Fr>(a).from_mont()); }
}
