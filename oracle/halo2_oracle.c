/* CPU restatement, in plain C, of the halo2_proofs / halo2curves hot path on BN254.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library
 * (anon-aadhaar-halo2_b200/) never links or calls it.
 *
 * PARITY PIN: per-function output vectors do not exist — the reference repository has no call
 * site, test or golden vector for this path and its Rust
 * dependencies cannot be built here (no cargo, no network).  End to end, the commitments this
 * file's best_multiexp produces are pinned by the reference's Solidity verifier accepting the
 * proofs that carry them (oracle/sol_verifier.py, tests/test_square_proof_oracle.py).  This file
 * restates the published algorithm of
 *   [DEP] halo2_proofs 0.2.0 @ v2023_01_20 (reference Cargo.lock:469-471)
 *         src/arithmetic.rs      best_fft, recursive_butterfly_arithmetic,
 *                                best_multiexp, multiexp_serial
 *         src/poly/domain.rs     ifft, distribute_powers_zeta, coeff_to_extended,
 *                                extended_to_coeff, divide_by_vanishing_poly
 *   [DEP] halo2curves 0.3.1 @ 0.3.1 (reference Cargo.lock:484-486)
 *         src/bn256/{fr,fq}.rs   4 x u64 Montgomery fields (R = 2^256)
 *         src/bn256/curve.rs     G1 Jacobian add / mixed add / double
 * and is validated against oracle/halo2_cpu.py + oracle/bn254.py (Python big integers,
 * mathematical definitions) in tests/test_oracle.py.  It doubles as the "restated CPU
 * baseline": the thread split mirrors rayon's (len / threads chunks for the MSM,
 * halving recursion for the FFT).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;           /* Montgomery form, fully reduced */
typedef struct { const uint64_t p[4]; uint64_t inv; fe one; fe r2; } field;

/* moduli: reference solidity_verifier_contract/contract.sol:210-211 */
static const field FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};
static const field FQ = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq_p(const uint64_t* a, const uint64_t* p) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > p[i]) return 1;
        if (a[i] < p[i]) return 0;
    }
    return 1;
}
static inline void sub_p(uint64_t* a, const uint64_t* p) {
    u128 b = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - p[i] - (uint64_t)b;
        a[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
static inline void f_add(const field* F, fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; ++i) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_p(t, F->p)) sub_p(t, F->p);   /* a + b < 2p < 2^255: no carry out */
    memcpy(r->l, t, 32);
}
static inline void f_sub(const field* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[4];
    u128 bw = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)bw;
        t[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)t[i] + F->p[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}
static inline void f_neg(const field* F, fe* r, const fe* a) {
    fe z = {{0, 0, 0, 0}};
    f_sub(F, r, &z, a);
}
/* Montgomery product, coarsely-integrated operand scanning (as halo2curves' field macro) */
static inline void f_mul(const field* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_p(t, F->p)) sub_p(t, F->p);
    memcpy(r->l, t, 32);
}
static inline void f_sqr(const field* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static inline void f_dbl(const field* F, fe* r, const fe* a) { f_add(F, r, a, a); }
static void f_pow(const field* F, fe* r, const fe* a, const uint64_t e[4]) {
    fe acc = F->one;
    for (int i = 255; i >= 0; --i) {
        f_sqr(F, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
    }
    *r = acc;
}
static void f_inv(const field* F, fe* r, const fe* a) {
    uint64_t e[4] = {F->p[0] - 2, F->p[1], F->p[2], F->p[3]};
    f_pow(F, r, a, e);
}
static void f_from_mont(const field* F, uint64_t out[4], const fe* a) {
    fe one = {{1, 0, 0, 0}}, t;
    f_mul(F, &t, a, &one);
    memcpy(out, t.l, 32);
}

/* ------------------------------------------------------------------ exported field ops */
void orc_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* r) { f_mul(&FR, (fe*)r, (const fe*)a, (const fe*)b); }
void orc_fq_mul(const uint64_t* a, const uint64_t* b, uint64_t* r) { f_mul(&FQ, (fe*)r, (const fe*)a, (const fe*)b); }
void orc_fr_inv(const uint64_t* a, uint64_t* r) { f_inv(&FR, (fe*)r, (const fe*)a); }

/* -------------------------------------------------------------------------------- G1 */
typedef struct { fe x, y; } g1a;        /* affine; identity = (0,0) */
typedef struct { fe x, y, z; } g1j;     /* Jacobian; identity z = 0 */

static inline int g1a_is_id(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1j_set_id(g1j* r) {
    memset(r, 0, sizeof *r);
    r->y = FQ.one;
}
/* dbl-2009-l (a = 0) */
static void g1j_double(g1j* r, const g1j* p) {
    if (fe_is_zero(&p->z)) { *r = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    f_sqr(&FQ, &a, &p->x);
    f_sqr(&FQ, &b, &p->y);
    f_sqr(&FQ, &c, &b);
    f_add(&FQ, &d, &p->x, &b);
    f_sqr(&FQ, &d, &d);
    f_sub(&FQ, &d, &d, &a);
    f_sub(&FQ, &d, &d, &c);
    f_dbl(&FQ, &d, &d);
    f_dbl(&FQ, &e, &a);
    f_add(&FQ, &e, &e, &a);
    f_sqr(&FQ, &f, &e);
    f_mul(&FQ, &z3, &p->z, &p->y);
    f_dbl(&FQ, &z3, &z3);
    f_dbl(&FQ, &t, &d);
    f_sub(&FQ, &x3, &f, &t);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_dbl(&FQ, &c, &c);
    f_sub(&FQ, &t, &d, &x3);
    f_mul(&FQ, &y3, &e, &t);
    f_sub(&FQ, &y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
/* madd-2007-bl with exceptional cases */
static void g1j_add_affine(g1j* r, const g1j* p, const g1a* q) {
    if (g1a_is_id(q)) { *r = *p; return; }
    if (fe_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = FQ.one; return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) { g1j_double(r, p); return; }
        g1j_set_id(r);
        return;
    }
    f_sub(&FQ, &h, &u2, &p->x);
    f_sqr(&FQ, &hh, &h);
    f_dbl(&FQ, &i, &hh);
    f_dbl(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &rr, &s2, &p->y);
    f_dbl(&FQ, &rr, &rr);
    f_mul(&FQ, &v, &p->x, &i);
    f_sqr(&FQ, &x3, &rr);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &rr, &t);
    f_mul(&FQ, &t, &p->y, &j);
    f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &h);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
/* add-2007-bl with exceptional cases */
static void g1j_add(g1j* r, const g1j* p, const g1j* q) {
    if (fe_is_zero(&p->z)) { *r = *q; return; }
    if (fe_is_zero(&q->z)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_sqr(&FQ, &z2z2, &q->z);
    f_mul(&FQ, &u1, &p->x, &z2z2);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s1, &p->y, &q->z);
    f_mul(&FQ, &s1, &s1, &z2z2);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { g1j_double(r, p); return; }
        g1j_set_id(r);
        return;
    }
    f_sub(&FQ, &h, &u2, &u1);
    f_dbl(&FQ, &i, &h);
    f_sqr(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &rr, &s2, &s1);
    f_dbl(&FQ, &rr, &rr);
    f_mul(&FQ, &v, &u1, &i);
    f_sqr(&FQ, &x3, &rr);
    f_sub(&FQ, &x3, &x3, &j);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &x3, &x3, &v);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &rr, &t);
    f_mul(&FQ, &t, &s1, &j);
    f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &q->z);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &z2z2);
    f_mul(&FQ, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_to_affine(g1a* r, const g1j* p) {
    if (fe_is_zero(&p->z)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    f_inv(&FQ, &zi, &p->z);
    f_sqr(&FQ, &zi2, &zi);
    f_mul(&FQ, &zi3, &zi2, &zi);
    f_mul(&FQ, &r->x, &p->x, &zi2);
    f_mul(&FQ, &r->y, &p->y, &zi3);
}
void orc_g1_to_affine(const uint64_t* jac12, uint64_t* aff8) { g1j_to_affine((g1a*)aff8, (const g1j*)jac12); }

/* ------------------------------------------------------------------------------- MSM */
/* arithmetic.rs multiexp_serial: unsigned c-bit windows, c = 1 | 3 | ceil(ln n),
 * 256/c + 1 segments from the top, lazily-typed buckets, running-sum reduction. */
static unsigned get_at(unsigned seg, unsigned c, const uint8_t* bytes) {
    unsigned skip_bits = seg * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    for (unsigned i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
    uint64_t tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    return (unsigned)(tmp % ((uint64_t)1 << c));
}
typedef struct { int kind; g1a a; g1j p; } bucket;  /* 0 none, 1 affine, 2 projective */

static void multiexp_serial(const fe* coeffs, const g1a* bases, size_t n, g1j* acc) {
    uint8_t* reprs = (uint8_t*)malloc(n * 32 + 8);
    for (size_t i = 0; i < n; ++i) f_from_mont(&FR, (uint64_t*)(reprs + 32 * i), &coeffs[i]);
    unsigned c = n < 4 ? 1 : (n < 32 ? 3 : (unsigned)ceil(log((double)n)));
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    bucket* buckets = (bucket*)malloc(nb * sizeof(bucket));
    for (int seg = (int)segments - 1; seg >= 0; --seg) {
        for (unsigned i = 0; i < c; ++i) g1j_double(acc, acc);
        for (size_t b = 0; b < nb; ++b) buckets[b].kind = 0;
        for (size_t i = 0; i < n; ++i) {
            unsigned d = get_at((unsigned)seg, c, reprs + 32 * i);
            if (!d) continue;
            bucket* B = &buckets[d - 1];
            if (B->kind == 0) { B->kind = 1; B->a = bases[i]; }
            else if (B->kind == 1) {
                g1j t;
                if (g1a_is_id(&B->a)) g1j_set_id(&t);
                else { t.x = B->a.x; t.y = B->a.y; t.z = FQ.one; }
                g1j_add_affine(&B->p, &t, &bases[i]);
                B->kind = 2;
            } else g1j_add_affine(&B->p, &B->p, &bases[i]);
        }
        g1j running;
        g1j_set_id(&running);
        for (size_t b = nb; b-- > 0;) {
            if (buckets[b].kind == 1) g1j_add_affine(&running, &running, &buckets[b].a);
            else if (buckets[b].kind == 2) g1j_add(&running, &running, &buckets[b].p);
            g1j_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(reprs);
}

typedef struct { const fe* c; const g1a* b; size_t n; g1j acc; } msm_job;
static void* msm_thread(void* arg) {
    msm_job* j = (msm_job*)arg;
    g1j_set_id(&j->acc);
    multiexp_serial(j->c, j->b, j->n, &j->acc);
    return NULL;
}
/* arithmetic.rs best_multiexp: chunks of len / threads (one extra short chunk when it
 * does not divide), one thread each, results folded in order. */
int orc_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t n, int threads, uint64_t* out12) {
    g1j total;
    g1j_set_id(&total);
    if (threads < 1) threads = 1;
    if (n > (size_t)threads) {
        size_t chunk = n / (size_t)threads;
        size_t nchunks = (n + chunk - 1) / chunk;
        msm_job* jobs = (msm_job*)malloc(nchunks * sizeof(msm_job));
        pthread_t* th = (pthread_t*)malloc(nchunks * sizeof(pthread_t));
        for (size_t k = 0; k < nchunks; ++k) {
            size_t s = k * chunk, len = (s + chunk <= n) ? chunk : n - s;
            jobs[k].c = (const fe*)coeffs + s;
            jobs[k].b = (const g1a*)bases + s;
            jobs[k].n = len;
            pthread_create(&th[k], NULL, msm_thread, &jobs[k]);
        }
        for (size_t k = 0; k < nchunks; ++k) {
            pthread_join(th[k], NULL);
            g1j_add(&total, &total, &jobs[k].acc);
        }
        free(jobs);
        free(th);
    } else {
        multiexp_serial((const fe*)coeffs, (const g1a*)bases, n, &total);
    }
    memcpy(out12, &total, sizeof total);
    return 0;
}

/* ------------------------------------------------------------------------------- FFT */
typedef struct { fe* a; size_t n; size_t tchunk; const fe* tw; int depth; } fft_job;
static void butterfly_layer(fe* a, size_t n, size_t tchunk, const fe* tw) {
    size_t half = n / 2;
    fe t = a[half];                       /* twiddle factor one */
    f_sub(&FR, &a[half], &a[0], &t);
    f_add(&FR, &a[0], &a[0], &t);
    for (size_t i = 1; i < half; ++i) {
        f_mul(&FR, &t, &a[half + i], &tw[i * tchunk]);
        fe u = a[i];
        f_add(&FR, &a[i], &u, &t);
        f_sub(&FR, &a[half + i], &u, &t);
    }
}
static void* fft_rec(void* arg);
/* arithmetic.rs recursive_butterfly_arithmetic: halves joined (rayon::join), then one
 * butterfly layer over the whole slice. */
static void recursive_butterfly(fe* a, size_t n, size_t tchunk, const fe* tw, int depth) {
    if (n == 2) {
        fe t = a[1];
        f_sub(&FR, &a[1], &a[0], &t);
        f_add(&FR, &a[0], &a[0], &t);
        return;
    }
    if (depth > 0 && n >= 4096) {
        fft_job j = {a + n / 2, n / 2, tchunk * 2, tw, depth - 1};
        pthread_t th;
        pthread_create(&th, NULL, fft_rec, &j);
        recursive_butterfly(a, n / 2, tchunk * 2, tw, depth - 1);
        pthread_join(th, NULL);
    } else {
        recursive_butterfly(a, n / 2, tchunk * 2, tw, 0);
        recursive_butterfly(a + n / 2, n / 2, tchunk * 2, tw, 0);
    }
    butterfly_layer(a, n, tchunk, tw);
}
static void* fft_rec(void* arg) {
    fft_job* j = (fft_job*)arg;
    recursive_butterfly(j->a, j->n, j->tchunk, j->tw, j->depth);
    return NULL;
}
static size_t bitreverse(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
/* arithmetic.rs best_fft: bit-reverse, n/2 twiddles, radix-2 DIT.  In place. */
int orc_best_fft(uint64_t* a_, const uint64_t* omega_, unsigned log_n, int threads) {
    fe* a = (fe*)a_;
    const fe* omega = (const fe*)omega_;
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return 0;
    for (size_t k = 0; k < n; ++k) {
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    size_t nt = n / 2 ? n / 2 : 1;
    fe* tw = (fe*)malloc(nt * sizeof(fe));
    tw[0] = FR.one;
    for (size_t i = 1; i < n / 2; ++i) f_mul(&FR, &tw[i], &tw[i - 1], omega);
    int depth = 0;
    while ((1 << (depth + 1)) <= threads) ++depth;
    recursive_butterfly(a, n, 1, tw, depth);
    free(tw);
    return 0;
}

typedef struct { fe* a; size_t lo, hi; const fe* tab; size_t tablen; int mode; } scale_job;
static void* scale_thread(void* arg) {
    scale_job* j = (scale_job*)arg;
    for (size_t i = j->lo; i < j->hi; ++i) {
        if (j->mode == 0) f_mul(&FR, &j->a[i], &j->a[i], &j->tab[0]);            /* all by tab[0] */
        else if (j->mode == 1) { size_t r = i % 3; if (r) f_mul(&FR, &j->a[i], &j->a[i], &j->tab[r - 1]); }
        else f_mul(&FR, &j->a[i], &j->a[i], &j->tab[i % j->tablen]);
    }
    return NULL;
}
static void par_scale(fe* a, size_t n, const fe* tab, size_t tablen, int mode, int threads) {
    if (threads < 1) threads = 1;
    if (n < 4096) threads = 1;
    pthread_t th[64];
    scale_job jobs[64];
    if (threads > 64) threads = 64;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo > n) lo = n;
        jobs[t] = (scale_job){a, lo, hi, tab, tablen, mode};
        pthread_create(&th[t], NULL, scale_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

/* domain.rs ifft: best_fft(omega_inv) then multiply by the divisor */
int orc_ifft(uint64_t* a, const uint64_t* omega_inv, unsigned log_n, const uint64_t* divisor, int threads) {
    orc_best_fft(a, omega_inv, log_n, threads);
    par_scale((fe*)a, (size_t)1 << log_n, (const fe*)divisor, 1, 0, threads);
    return 0;
}
/* domain.rs coeff_to_extended: distribute_powers_zeta(into_coset), resize, best_fft */
int orc_coeff_to_extended(const uint64_t* in, unsigned k, uint64_t* out, unsigned ext_k, const uint64_t* ext_omega,
                          const uint64_t* zeta, int threads) {
    size_t n = (size_t)1 << k, N = (size_t)1 << ext_k;
    fe tab[2];
    tab[0] = *(const fe*)zeta;
    f_sqr(&FR, &tab[1], &tab[0]);          /* g_coset_inv = zeta^2 */
    memcpy(out, in, n * 32);
    memset(out + 4 * n, 0, (N - n) * 32);
    par_scale((fe*)out, n, tab, 2, 1, threads);
    return orc_best_fft(out, ext_omega, ext_k, threads);
}
/* domain.rs extended_to_coeff: ifft, distribute_powers_zeta(out of coset), truncate */
int orc_extended_to_coeff(uint64_t* a, unsigned ext_k, const uint64_t* ext_omega_inv, const uint64_t* divisor,
                          const uint64_t* zeta, int threads) {
    size_t N = (size_t)1 << ext_k;
    orc_ifft(a, ext_omega_inv, ext_k, divisor, threads);
    fe tab[2];
    f_sqr(&FR, &tab[0], (const fe*)zeta);  /* [g_coset_inv, g_coset] */
    tab[1] = *(const fe*)zeta;
    par_scale((fe*)a, N, tab, 2, 1, threads);
    return 0;  /* caller truncates to n * (d - 1) */
}
/* domain.rs divide_by_vanishing_poly */
int orc_divide_by_vanishing(uint64_t* h, unsigned ext_k, const uint64_t* t_eval, size_t t_len, int threads) {
    par_scale((fe*)h, (size_t)1 << ext_k, (const fe*)t_eval, t_len, 2, threads);
    return 0;
}

/* --------------------------------------------------------------------- seeded inputs */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* oracle/bn254.py seeded_fr_mont_limbs */
void orc_gen_scalars(uint64_t* out, size_t n, uint64_t seed, size_t start) {
    for (size_t i = 0; i < n; ++i) {
        uint64_t* l = out + 4 * i;
        for (int j = 0; j < 4; ++j) l[j] = splitmix64((seed << 32) + 4 * (start + i) + (uint64_t)j);
        l[3] &= ((uint64_t)1 << 62) - 1;
        if (geq_p(l, FR.p)) sub_p(l, FR.p);
    }
}
/* P_i = [t_i] G with t_i = splitmix64(seed<<32 + i) | 1: byte-window tables tbl[j][b] = [b * 2^(8j)] G
 * (8 mixed additions per point) and one field inversion per 256 points (Montgomery's trick) for the
 * affine form.  The points are unique group elements in unique affine form, so the table layout does
 * not show in the output (tests/test_oracle.py pins them against oracle/bn254.py). */
typedef struct { uint64_t* out; size_t lo, hi, start; uint64_t seed; const g1a* tbl; } gen_job;
#define GEN_BATCH 256
static void* gen_points_thread(void* arg) {
    gen_job* j = (gen_job*)arg;
    g1j acc[GEN_BATCH];
    fe pre[GEN_BATCH];
    for (size_t i0 = j->lo; i0 < j->hi; i0 += GEN_BATCH) {
        size_t m = j->hi - i0 < GEN_BATCH ? j->hi - i0 : GEN_BATCH;
        fe run = FQ.one;
        for (size_t u = 0; u < m; ++u) {
            uint64_t t = splitmix64((j->seed << 32) + (j->start + i0 + u)) | 1ull;
            g1j_set_id(&acc[u]);
            for (int b = 0; b < 8; ++b) {
                unsigned d = (unsigned)(t >> (8 * b)) & 255u;
                if (d) g1j_add_affine(&acc[u], &acc[u], &j->tbl[b * 256 + d]);
            }
            pre[u] = run;                      /* product of the z of the earlier points of the batch */
            if (!fe_is_zero(&acc[u].z)) f_mul(&FQ, &run, &run, &acc[u].z);
        }
        fe inv;
        f_inv(&FQ, &inv, &run);
        for (size_t u = m; u-- > 0;) {
            g1a* o = (g1a*)(j->out + 8 * (i0 + u));
            if (fe_is_zero(&acc[u].z)) { memset(o, 0, sizeof *o); continue; }
            fe zi, zi2, zi3;
            f_mul(&FQ, &zi, &inv, &pre[u]);
            f_mul(&FQ, &inv, &inv, &acc[u].z);
            f_sqr(&FQ, &zi2, &zi);
            f_mul(&FQ, &zi3, &zi2, &zi);
            f_mul(&FQ, &o->x, &acc[u].x, &zi2);
            f_mul(&FQ, &o->y, &acc[u].y, &zi3);
        }
    }
    return NULL;
}
/* oracle/bn254.py seeded_g1_points: P_i = [splitmix64(seed<<32 + i) | 1] G, G = (1, 2) */
void orc_gen_points(uint64_t* out, size_t n, uint64_t seed, size_t start, int threads) {
    static g1a tbl[8 * 256];
    static int tbl_ready = 0;
    static pthread_mutex_t tbl_mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&tbl_mu);
    if (!tbl_ready) {
        g1j base;
        base.x = FQ.one;
        f_dbl(&FQ, &base.y, &FQ.one);
        base.z = FQ.one;
        for (int b = 0; b < 8; ++b) {
            g1a step;
            g1j_to_affine(&step, &base);       /* [2^(8b)] G */
            g1j acc;
            g1j_set_id(&acc);
            memset(&tbl[b * 256], 0, sizeof(g1a));
            for (int d = 1; d < 256; ++d) {
                g1j_add_affine(&acc, &acc, &step);
                g1j_to_affine(&tbl[b * 256 + d], &acc);
            }
            for (int q = 0; q < 8; ++q) g1j_double(&base, &base);
        }
        tbl_ready = 1;
    }
    pthread_mutex_unlock(&tbl_mu);
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    gen_job jobs[64];
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo > n) lo = n;
        jobs[t] = (gen_job){out, lo, hi, start, seed, tbl};
        pthread_create(&th[t], NULL, gen_points_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------- local KZG setup (test SRS) */
/* out[i] = [scalars[i]] G, G = (1, 2): scalars Montgomery Fr, points affine Montgomery.
 * The group elements a `ParamsKZG::setup(k, rng)` produces ([s^i] G and [L_i(s)] G,
 * poly/kzg/commitment.rs) are determined by the scalars; this is plain double-and-add over a
 * table of [2^b] G. */
typedef struct { const uint64_t* sc; uint64_t* out; size_t lo, hi; const g1a* tbl; } gmul_job;
static void* gmul_thread(void* arg) {
    gmul_job* j = (gmul_job*)arg;
    for (size_t i = j->lo; i < j->hi; ++i) {
        uint64_t s[4];
        f_from_mont(&FR, s, (const fe*)(j->sc + 4 * i));
        g1j acc;
        g1j_set_id(&acc);
        for (int b = 0; b < 254; ++b)
            if ((s[b >> 6] >> (b & 63)) & 1) g1j_add_affine(&acc, &acc, &j->tbl[b]);
        g1j_to_affine((g1a*)(j->out + 8 * i), &acc);
    }
    return NULL;
}
void orc_g1_generator_mul(const uint64_t* scalars, size_t n, uint64_t* out, int threads) {
    static g1a tbl[254];
    g1j p;
    p.x = FQ.one;
    f_dbl(&FQ, &p.y, &FQ.one);
    p.z = FQ.one;
    for (int b = 0; b < 254; ++b) {
        g1j_to_affine(&tbl[b], &p);
        g1j_double(&p, &p);
    }
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    gmul_job jobs[64];
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo > n) lo = n;
        jobs[t] = (gmul_job){scalars, out, lo, hi, tbl};
        pthread_create(&th[t], NULL, gmul_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
