/* b200zk — C ABI of the B200-native Halo2/KZG prover hot path (BN254).
 *
 * This is the drop-in boundary a patched `halo2_proofs` binds over FFI.  The
 * reference repository (anon-aadhaar/anon-aadhaar-halo2) contains no prover code: the
 * functions replaced here live in its pinned, un-vendored dependency
 *   halo2_proofs 0.2.0, tag v2023_01_20, rev c7e42e41   (reference Cargo.lock:469-471)
 *   halo2curves  0.3.1, tag 0.3.1,       rev 9b67e19b   (reference Cargo.lock:484-486)
 * and are reached from the reference only through `halo2_base::halo2_proofs`
 * (reference src/lib.rs:15-18).  Each entry point below names the upstream Rust item
 * whose body it replaces; INTEGRATION.md shows the Rust-side `extern "C"` block.
 *
 * Conventions (SURVEY.md section 8 b)
 *   - Field elements: 4 little-endian u64 limbs, Montgomery form (x * 2^256 mod p),
 *     fully reduced, exactly the in-memory layout of `bn256::Fr` / `bn256::Fq`.
 *   - G1Affine: x limbs then y limbs (64 bytes); the identity is all-zero.
 *   - G1 (return type of best_multiexp): Jacobian x, y, z (96 bytes); identity = (0, 1, 0).
 *   - Every function returns 0 on success, non-zero on failure; the message is
 *     available from b200zk_last_error().  Upstream's functions are infallible and
 *     assert on bad lengths, so the Rust shim panics with that message.
 *   - Host pointers are never retained past the call.  `*_dev` entry points take device
 *     pointers and a CUDA stream and are asynchronous on that stream.
 *   - No CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifetime ---------------------------------------------------------------------- */
/* Bind the library to CUDA device `device` (-1: current device).  Idempotent. */
int b200zk_init(int device);
int b200zk_shutdown(void);
const char* b200zk_last_error(void);
/* ABI version of this header (checked by the bindings). */
uint32_t b200zk_abi_version(void);

/* ---- NTT: arithmetic::best_fft and EvaluationDomain -------------------------------- */
/* halo2_proofs/src/arithmetic.rs `best_fft::<Fr>(a, omega, log_n)`: in place,
 * natural order in and out, a'[i] = sum_j a[j] * omega^(i*j). */
int b200zk_ntt(uint64_t* a, uint32_t log_n, const uint64_t omega[4]);

/* halo2_proofs/src/poly/domain.rs `EvaluationDomain::ifft(a, omega_inv, log_n, divisor)`
 * as used by `lagrange_to_coeff`: best_fft with omega_inv, then a[i] *= divisor. */
int b200zk_intt(uint64_t* a, uint32_t log_n, const uint64_t omega_inv[4], const uint64_t divisor[4]);

/* `EvaluationDomain::coeff_to_extended`: in[i] *= zeta^(i mod 3), zero-pad 2^k -> 2^ext_k,
 * best_fft with extended_omega.  `in` has 2^k elements, `out` 2^ext_k (may not alias). */
int b200zk_coeff_to_extended(const uint64_t* in, uint32_t k, uint64_t* out, uint32_t ext_k,
                             const uint64_t extended_omega[4], const uint64_t zeta[4]);

/* `EvaluationDomain::extended_to_coeff`: ifft with extended_omega_inv and
 * extended_ifft_divisor, a[i] *= zeta^-(i mod 3), truncate to `keep` = n * (d - 1)
 * elements.  `a` has 2^ext_k elements; the first `keep` are written to `out`. */
int b200zk_extended_to_coeff(const uint64_t* a, uint32_t ext_k, const uint64_t extended_omega_inv[4],
                             const uint64_t extended_ifft_divisor[4], const uint64_t zeta[4], uint64_t* out,
                             size_t keep);

/* `EvaluationDomain::divide_by_vanishing_poly`: h[i] *= t_evaluations[i mod t_len] in
 * place; t_evaluations are the (already inverted) values upstream stores, t_len = 2^(ext_k - k). */
int b200zk_divide_by_vanishing(uint64_t* h, uint32_t ext_k, const uint64_t* t_evaluations, uint32_t t_len);

/* Batched forms: `count` independent transforms, element `c` of the batch at
 * `a + c * stride` (stride in field elements).  Same semantics per column. */
int b200zk_ntt_many(uint64_t* a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega[4]);
int b200zk_intt_many(uint64_t* a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega_inv[4],
                     const uint64_t divisor[4]);
int b200zk_coeff_to_extended_many(const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride,
                                  size_t count, uint32_t k, uint32_t ext_k, const uint64_t extended_omega[4],
                                  const uint64_t zeta[4]);

/* Device-resident forms (pointers are CUDA device pointers; `stream` is a cudaStream_t,
 * NULL = the library's stream).  `d_a` is transformed in place. */
int b200zk_ntt_dev(void* d_a, size_t stride, size_t count, uint32_t log_n, const uint64_t omega[4],
                   const uint64_t* divisor_or_null, void* stream);
int b200zk_coeff_to_extended_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                                 uint32_t k, uint32_t ext_k, const uint64_t extended_omega[4],
                                 const uint64_t zeta[4], void* stream);
int b200zk_extended_to_coeff_dev(const void* d_a, uint32_t ext_k, const uint64_t extended_omega_inv[4],
                                 const uint64_t extended_ifft_divisor[4], const uint64_t zeta[4],
                                 const void* d_t_evaluations_or_null, uint32_t t_len, void* d_out, size_t keep,
                                 void* stream);

/* ---- MSM: arithmetic::best_multiexp and ParamsKZG::commit / commit_lagrange -------- */
/* halo2_proofs/src/arithmetic.rs `best_multiexp::<G1Affine>(coeffs, bases) -> G1`.
 * scalars: n x 4 limbs (Fr Montgomery); bases: n x 8 limbs (G1Affine); out: 12 limbs (G1). */
int b200zk_msm_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_xyz[12]);

/* halo2_proofs/src/poly/kzg/commitment.rs: `ParamsKZG::g` / `g_lagrange` are fixed for the
 * life of the params, so they are uploaded once and addressed by handle.
 * `commit(poly)` == msm_g1_registered(handle(g), poly, len);
 * `commit_lagrange(poly)` == msm_g1_registered(handle(g_lagrange), poly, len). */
int b200zk_bases_register(const uint64_t* bases, size_t n, uint64_t* handle_out);
int b200zk_bases_evict(uint64_t handle);
int b200zk_msm_g1_registered(uint64_t handle, const uint64_t* scalars, size_t n, uint64_t out_xyz[12]);

/* Device-resident form: d_scalars / d_bases are device pointers; result (12 limbs) is
 * written to host memory `out_xyz` after the stream is synchronised. */
int b200zk_msm_g1_dev(const void* d_scalars, const void* d_bases, size_t n, uint64_t out_xyz[12], void* stream);
/* Same, result left in device memory (12 limbs at d_out), fully asynchronous. */
int b200zk_msm_g1_dev_async(const void* d_scalars, const void* d_bases, size_t n, void* d_out_xyz, void* stream);

/* Sum of `count` Jacobian points (12 limbs each, host memory) -> out: the fold
 * best_multiexp applies to its per-chunk results; used to combine per-GPU partial MSMs. */
int b200zk_g1_sum(const uint64_t* points_xyz, size_t count, uint64_t out_xyz[12]);

/* ---- synthetic inputs (benchmark / test support; oracle/bn254.py defines the streams) -- */
int b200zk_gen_scalars_dev(void* d_out, size_t n, uint64_t seed, size_t start);
int b200zk_gen_points_dev(void* d_out, size_t n, uint64_t seed, size_t start);

/* ---- measurement support ------------------------------------------------------------ */
/* Register-resident Fq multiply chain on every SM; returns field multiplications / s.
 * This is the measured denominator of the MSM integer-pipe roofline. */
int b200zk_modmul_peak(uint32_t iters, double* modmul_per_s_out);
/* Per-stage CUDA-event timing of the MSM pipeline (recorded on the stream the kernels run
 * on).  After a profiled MSM, b200zk_msm_last_stages fills ms_out[0..8] with the stage
 * durations {hist, scan, scatter, sync, accumulate, combine, reduce, reduce-combine, fold}
 * and info_out with {n, window bits c, windows, (bucket, point) pairs, chunk length}. */
int b200zk_msm_profile(int enable);
int b200zk_msm_last_stages(float* ms_out, int capacity, uint64_t info_out[5]);
/* Number of kernels launched by this library since init (for bench.py gpu_launches). */
uint64_t b200zk_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
