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
 *   - Host pointers are never retained past the call (small host tables a `*_dev` call reads are
 *     copied into library-owned page-locked slots before it returns).  `*_dev` entry points take
 *     device pointers and a CUDA stream and are asynchronous on that stream: they return without
 *     waiting for the device.  Calls are serialised on the host by one mutex, but the work they
 *     queue on different streams runs concurrently: every stream the library is called on gets its
 *     own scratch space (b200zk_stream_release frees it), so two MSMs / transforms on two streams
 *     do not share any work buffer.  Ordering between streams is the caller's (events).
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
/* Free the scratch space the library keeps for `stream` (NULL = its own stream); call it before
 * destroying a stream that was passed to `*_dev` entry points.  Synchronises the device. */
int b200zk_stream_release(void* stream);

/* ---- host memory ------------------------------------------------------------------- */
/* Every host-pointer entry point accepts pageable memory (a Rust `Vec<Fr>`), but the driver then
 * stages each transfer through its own bounce buffer at a fraction of the PCIe rate and the
 * upload pipeline of b200zk_msm_g1_registered cannot overlap it.  A shim that keeps a buffer for
 * the whole proof (the polynomials of `create_proof`, the SRS during registration) page-locks it
 * once with b200zk_host_register; a caller that chooses its allocator takes page-locked memory
 * from b200zk_host_alloc.  Registration does not transfer ownership: the caller unregisters
 * before freeing.  Everything still registered or allocated is released by b200zk_shutdown. */
int b200zk_host_register(void* ptr, size_t bytes);
int b200zk_host_unregister(void* ptr);
int b200zk_host_alloc(size_t bytes, void** ptr_out);
int b200zk_host_free(void* ptr);

/* Device mirrors of host polynomials (off by default).  `create_proof` hands the same `Polynomial`
 * to several of the calls below in a row — commit_lagrange(p), lagrange_to_coeff(p),
 * coeff_to_extended(&p), evaluate_h(.., p, ..) — and every call would upload it again.  With
 * b200zk_mirror_enable(max_bytes > 0) each host Fr buffer a host-pointer call reads is kept in HBM,
 * keyed by (pointer, length), and a call that writes a host buffer (the in-place transforms, the
 * `out` of coeff_to_extended / extended_to_coeff, b200zk_dev_download) leaves its mirror equal to
 * what it wrote; later calls on the same buffer skip the upload (b200zk_dev_upload copies device to
 * device).  Results are the same bytes either way.  The contract that makes this sound: while mirrors
 * are on, the caller calls b200zk_mirror_invalidate(ptr, bytes) before it writes or frees host memory
 * it has passed to a host-pointer call (bytes = 0: whatever mirror contains `ptr`; ptr = NULL: every
 * mirror, e.g. between proofs — the device blocks stay pooled and the statistics restart).  The Rust fork
 * keeps that contract inside `Polynomial` (a `Cell<bool>` set when the buffer is handed to the
 * library, checked in `DerefMut` and `Drop`) and invalidates immediately after a call on a plain
 * slice (`best_multiexp`, `best_fft`), see INTEGRATION.md.  Least-recently-used mirrors are dropped
 * beyond max_bytes; max_bytes = 0 drops everything and turns mirroring off.
 * b200zk_mirror_stats: {hits, misses, resident bytes, evictions}. */
int b200zk_mirror_enable(size_t max_bytes);
int b200zk_mirror_invalidate(const void* host_ptr, size_t bytes);
int b200zk_mirror_stats(uint64_t out[4]);

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

/* Cosets of the extended domain, for evaluating h(X) on several GPUs (SURVEY.md section 8 e).  The
 * extended domain {zeta * w_ext^e} is the union of P = 2^(ext_k - k) cosets of the base domain:
 * e = P j + q is the point (zeta w_ext^q) w^j, and a rotation by w stays inside a coset.  So the
 * rank that owns coset q needs, per column, one 2^k-point transform of the coefficients scaled by
 * g^i with g = zeta * w_ext^q (`b200zk_coeff_to_coset_dev`), the matching slices of the proving
 * key's extended columns (`b200zk_extended_coset_slice_dev`), and runs the quotient kernels with
 * k = ext_k = k, extended_omega = w, zeta = g.  `b200zk_extended_coset_interleave_dev` puts an
 * evaluated coset back at e = P j + q of the extended numerator before extended_to_coeff.
 * Equivalent to coeff_to_extended followed by taking every P-th element, at 1/P of the work. */
int b200zk_coeff_to_coset_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                              uint32_t k, const uint64_t omega[4], const uint64_t coset_generator[4], void* stream);
int b200zk_extended_coset_slice_dev(const void* d_ext, size_t ext_stride, void* d_out, size_t out_stride, size_t count,
                                    uint32_t k, uint32_t ext_k, uint32_t coset, void* stream);
int b200zk_extended_coset_interleave_dev(const void* d_coset, void* d_ext, uint32_t k, uint32_t ext_k, uint32_t coset,
                                         void* stream);

/* One best_fft of 2^log_n elements sharded over `world` (1, 2, 4 or 8) GPUs, one process per
 * GPU (SURVEY.md section 8 e).  n = n1 * n2 with n1 = 2^log_n1 <= 2^9; rank r holds the columns
 * j2 in [r m, (r + 1) m), m = n2 / world, of the natural-order vector seen as an [n1][n2] matrix:
 * d_in[j1][j2 - r m] = a[j1 * n2 + j2].
 *   1. b200zk_ntt4_first_pass_scatter_dev — ONE kernel: the n1-point transforms along j1 (pass 0 of
 *      the whole transform, restricted to the rank's columns), the twiddle omega^(i1 * j2), and the
 *      exchange: element (i1, j2) is stored at dest_bases[i1 / (n1 / world)] +
 *      (i1 mod (n1 / world)) * dest_pitch + dest_col_offset + (j2 - r m)   [offsets in field
 *      elements].  With peer-mapped buffers (dest_pitch = n2, dest_col_offset = r m) the stores
 *      travel over NVLink and the pass is itself the all-to-all; with a local send buffer
 *      (dest_bases[s] = send + s * (n1 / world) * m, dest_pitch = m, offset 0) it packs for a NCCL
 *      all-to-all, whose result b200zk_ntt4_gather_rows_dev turns into the [n1 / world][n2] row block;
 *   2. b200zk_ntt_dev(rows, n2, n1 / world, log_n2, omega^n1): row i1 then holds A[i1 + n1 * i2] at
 *      position i2.
 * `dest_bases` is a host array of `world` device pointers.
 * b200zk_ntt4_twiddle_scatter_dev is the unfused form of step 1 for a column-major slab
 * (d_in[j2 - r m][j1], any n1): the caller first runs b200zk_ntt_dev(d_in, n1, m, log_n1, omega^n2),
 * then this kernel multiplies by omega^(i1 * j2) and stores exactly as above. */
int b200zk_ntt4_first_pass_scatter_dev(const void* d_in, uint32_t log_n, uint32_t log_n1, const uint64_t omega[4],
                                       uint32_t world, uint32_t rank, void* const* dest_bases, size_t dest_pitch,
                                       size_t dest_col_offset, void* stream);
int b200zk_ntt4_twiddle_scatter_dev(const void* d_in, uint32_t log_n, uint32_t log_n1, const uint64_t omega[4],
                                    uint32_t world, uint32_t rank, void* const* dest_bases, size_t dest_pitch,
                                    size_t dest_col_offset, void* stream);
int b200zk_ntt4_gather_rows_dev(const void* d_recv, void* d_out, uint32_t log_n, uint32_t log_n1, uint32_t world,
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
/* Same with control over the per-window table T[w][i] = 2^(c*w) * bases[i] that registration
 * builds by default (W x the memory of the bases; every window then shares one bucket set and
 * the final doubling chain disappears — the win for the prover's 2^15..2^17 commitments).
 * precompute_windows = 0 keeps only the points, 1 lets the cost model choose the window size,
 * 4..23 forces that many window bits. */
int b200zk_bases_register_ex(const uint64_t* bases, size_t n, int precompute_windows, uint64_t* handle_out);
int b200zk_bases_evict(uint64_t handle);
int b200zk_msm_g1_registered(uint64_t handle, const uint64_t* scalars, size_t n, uint64_t out_xyz[12]);
/* `count` commitments against the same registered bases in one pipeline: column j is
 * scalars + j * stride (n scalars each), result j at out_xyz + 12 * j.  This is the batched
 * entry point behind create_proof's `advice.iter().map(|p| params.commit_lagrange(p))`. */
int b200zk_msm_g1_registered_many(uint64_t handle, const uint64_t* scalars, size_t stride, size_t count, size_t n,
                                  uint64_t* out_xyz);
/* Device-resident scalars and results (count x 12 limbs at d_out_xyz), asynchronous on `stream`:
 * the pair count the sort produces stays on the device (it plans the accumulation there). */
int b200zk_msm_g1_registered_dev(uint64_t handle, const void* d_scalars, size_t stride, size_t count, size_t n,
                                 void* d_out_xyz, void* stream);

/* Device-resident form: d_scalars / d_bases are device pointers; result (12 limbs) is
 * written to host memory `out_xyz` after the stream is synchronised. */
int b200zk_msm_g1_dev(const void* d_scalars, const void* d_bases, size_t n, uint64_t out_xyz[12], void* stream);
/* Same, result left in device memory (12 limbs at d_out), fully asynchronous. */
int b200zk_msm_g1_dev_async(const void* d_scalars, const void* d_bases, size_t n, void* d_out_xyz, void* stream);

/* Sum of `count` Jacobian points (12 limbs each, host memory) -> out: the fold
 * best_multiexp applies to its per-chunk results; used to combine per-GPU partial MSMs. */
int b200zk_g1_sum(const uint64_t* points_xyz, size_t count, uint64_t out_xyz[12]);

/* ---- device-resident Fr columns -------------------------------------------------------- */
/* The quotient evaluation reads hundreds of extended columns; they live in HBM and are
 * addressed by handle.  Sizes and offsets are in field elements (32 bytes). */
int b200zk_dev_alloc(size_t n_elems, uint64_t* handle_out);
int b200zk_dev_free(uint64_t handle);
/* A second handle onto elements [offset, offset + n_elems) of `parent` (no copy; freeing the
 * view leaves the parent alone, the parent must outlive it).  Lets a batch of columns live in
 * one allocation for the strided `*_dev` transforms while the quotient kernels address them
 * one by one. */
int b200zk_dev_view(uint64_t parent, size_t offset, size_t n_elems, uint64_t* handle_out);
int b200zk_dev_upload(uint64_t handle, size_t offset, const uint64_t* host, size_t n_elems);
int b200zk_dev_download(uint64_t handle, size_t offset, uint64_t* host, size_t n_elems);
/* Raw device pointer of a handle (NULL if unknown), for the *_dev entry points above. */
void* b200zk_dev_ptr(uint64_t handle);

/* ---- quotient: plonk::evaluation::Evaluator::evaluate_h ------------------------------ */
/* Flat encoding of halo2_proofs/src/plonk/evaluation.rs `GraphEvaluator`.
 * b200zk_src mirrors `ValueSource`: kind 0 Constant(a) 1 Intermediate(a) 2 Fixed(a, rot b)
 * 3 Advice(a, rot b) 4 Instance(a, rot b) 5 Challenge(a) 6 Beta 7 Gamma 8 Theta 9 Y
 * 10 PreviousValue; `b` indexes `rotations`.
 * b200zk_calc mirrors `CalculationInfo`: op 0 Add(x,y) 1 Sub(x,y) 2 Mul(x,y) 3 Square(x)
 * 4 Double(x) 5 Negate(x) 6 Horner(start = x, parts[parts_off..+parts_len], factor = y)
 * 7 Store(x); the result goes to intermediates[target]. */
typedef struct { uint32_t kind, a, b; } b200zk_src;
typedef struct { uint32_t op, target; b200zk_src x, y; uint32_t parts_off, parts_len; } b200zk_calc;
typedef struct {
    const uint64_t* constants;  uint32_t n_constants;   /* n x 4 limbs */
    const int32_t* rotations;   uint32_t n_rotations;
    const b200zk_calc* calcs;   uint32_t n_calcs;
    const b200zk_src* parts;    uint32_t n_parts;
    uint32_t n_intermediates;
} b200zk_graph;
/* Everything `GraphEvaluator::evaluate` reads besides the graph: extended-domain column
 * handles (each 2^ext_k elements), challenges (n x 4 limbs) and the four scalars. */
typedef struct {
    const uint64_t* fixed;     uint32_t n_fixed;       /* handles */
    const uint64_t* advice;    uint32_t n_advice;
    const uint64_t* instance;  uint32_t n_instance;
    const uint64_t* challenges; uint32_t n_challenges;
    uint64_t beta[4], gamma[4], theta[4], y[4];
    uint32_t k, ext_k;
    /* Multi-GPU sharding of the extended domain: only indices [range_begin, range_begin +
     * range_len) are evaluated and written (range_len = 0: the whole domain).  Inputs stay
     * full columns because rotations reach outside the range. */
    uint64_t range_begin, range_len;
} b200zk_quotient_env;

/* The three quotient entry points are asynchronous on the library's stream (their results stay in
 * device columns); b200zk_dev_download, or any host-result call, synchronises.
 *
 * out[idx] = graph.evaluate(idx, previous_value = previous[idx] or 0 when the handle is 0)
 * for every idx of the extended domain; rotation r reads index
 * (idx + r * 2^(ext_k - k)) mod 2^ext_k.  `out` may equal `previous` (the custom-gate
 * loop of evaluate_h: values[idx] = custom_gates.evaluate(.., &values[idx], ..)). */
int b200zk_quotient_graph(const b200zk_graph* graph, const b200zk_quotient_env* env, uint64_t previous_handle,
                          uint64_t out_handle);

/* The permutation-argument block of evaluate_h, folded into `values` in upstream's order:
 * l_0 (1 - z_0); l_last (z_l^2 - z_l); l_0 (z_i - z_{i-1}(omega^last X)) for i > 0; then per
 * set (z_i(omega X) prod(v + beta sigma + gamma) - z_i(X) prod(v + delta^j beta X + gamma)) * l_active.
 * column_kind[j] in {2 fixed, 3 advice, 4 instance} and column_index[j] name permutation
 * column j inside `env`; sigma[j] is pk.permutation.cosets[j]; products[i] the z_i cosets. */
int b200zk_quotient_permutation(const b200zk_quotient_env* env, uint64_t values_handle,
                                const uint32_t* column_kind, const uint32_t* column_index,
                                const uint64_t* sigma_handles, uint32_t n_columns,
                                const uint64_t* product_handles, uint32_t n_sets, uint32_t chunk_len,
                                uint32_t blinding_factors, uint64_t l0_handle, uint64_t l_last_handle,
                                uint64_t l_active_row_handle, const uint64_t extended_omega[4],
                                const uint64_t zeta[4], const uint64_t delta[4]);

/* One lookup argument's five constraints folded into `values`; `table_values` is the
 * output of b200zk_quotient_graph on that lookup's graph with previous = 0. */
int b200zk_quotient_lookup(const b200zk_quotient_env* env, uint64_t values_handle, uint64_t table_values_handle,
                           uint64_t product_handle, uint64_t permuted_input_handle,
                           uint64_t permuted_table_handle, uint64_t l0_handle, uint64_t l_last_handle,
                           uint64_t l_active_row_handle);

/* ---- column arithmetic around the commitments (SURVEY.md section 8 f) ------------------- */
/* ff 0.12 `BatchInvert::batch_invert` (reference Cargo.lock:339-341): a[i] <- a[i]^-1 in place,
 * zeros stay zero. */
int b200zk_batch_invert(uint64_t* a, size_t n);
int b200zk_batch_invert_dev(void* d_a, size_t n, void* stream);

/* out[c][0] = first[c] (1 when null), out[c][i] = out[c][i-1] * in[c][i-1] for `count` columns of
 * n elements: the running products of the permutation and lookup arguments.  `first` is a host
 * array of count x 4 limbs.  `d_out` may equal `d_in`. */
int b200zk_prefix_product_dev(const void* d_in, size_t in_stride, void* d_out, size_t out_stride, size_t count,
                              size_t n, const uint64_t* first_or_null, void* stream);

/* halo2_proofs/src/plonk/permutation/prover.rs `Argument::commit`, the arithmetic of every set:
 * for chunks of `chunk_len` columns, modified[i] = prod_col (v_col[i] + delta^j beta omega^i + gamma)
 * / prod_col (v_col[i] + beta sigma_col[i] + gamma) (j = column index over the whole argument), then
 * z_s[0] = z_{s-1}[n - blinding_factors - 1] (1 for the first set), z_s[i+1] = z_s[i] * modified[i].
 * d_values / d_sigma: host arrays of n_cols device pointers to Lagrange-basis columns of 2^k
 * elements (column values and pkey.permutations).  d_z receives ceil(n_cols / chunk_len) columns
 * of 2^k elements.  Upstream overwrites the last `blinding_factors` rows of each z with
 * `Scalar::random(rng)`: pass those scalars (sets x blinding_factors x 4 limbs, host) in `blinds`
 * so the caller's RNG stream is the one consumed, or null to leave the computed values. */
int b200zk_permutation_product_dev(const void* const* d_values, const void* const* d_sigma, uint32_t n_cols,
                                   uint32_t chunk_len, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4],
                                   const uint64_t omega[4], const uint64_t delta[4], uint32_t blinding_factors,
                                   const uint64_t* blinds_or_null, void* d_z, void* stream);

/* halo2_proofs/src/plonk/lookup/prover.rs `Permuted::commit_product` for `count` lookups:
 * z[0] = 1, z[i+1] = z[i] * (compressed_input[i] + beta)(compressed_table[i] + gamma)
 *                         / ((permuted_input[i] + beta)(permuted_table[i] + gamma)),
 * last `blinding_factors` rows from `blinds` (count x blinding_factors x 4 limbs) when given. */
int b200zk_lookup_product_dev(const void* const* d_compressed_input, const void* const* d_compressed_table,
                              const void* const* d_permuted_input, const void* const* d_permuted_table,
                              uint32_t count, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4],
                              uint32_t blinding_factors, const uint64_t* blinds_or_null, void* d_z, void* stream);

/* halo2_proofs/src/plonk/lookup/prover.rs `permute_expression_pair` for `count` lookups: over the
 * usable rows m = 2^k - (blinding_factors + 1), permuted_input = the input expression sorted by
 * Fr's Ord (canonical integers); permuted_table carries each value at the row where it first
 * appears in permuted_input, and the table's remaining values, ascending, at the rows of repeated
 * inputs from the last such row backwards.  The last blinding_factors + 1 rows come from `blinds`
 * (host; per lookup the input's scalars then the table's, count x 2 x (blinding_factors + 1) x 4
 * limbs; zeros when null).  Column c of every array is at + c * stride elements.  Fails with
 * "ConstraintSystemFailure" in the message when an input value does not occur in the table. */
int b200zk_permute_expression_pair_dev(const void* d_input, const void* d_table, size_t stride, uint32_t count, uint32_t k,
                                       uint32_t blinding_factors, const uint64_t* blinds_or_null, void* d_permuted_input,
                                       void* d_permuted_table, size_t out_stride, void* stream);

/* halo2_proofs/src/arithmetic.rs `eval_polynomial(poly, point)` for `count` polynomials of n
 * coefficients (polynomial c at d_polys + c * stride), each at its own point (host, count x 4
 * limbs); results to host memory `out` (count x 4 limbs). */
int b200zk_eval_polynomial_dev(const void* d_polys, size_t stride, size_t count, size_t n, const uint64_t* points,
                               uint64_t* out, void* stream);

/* halo2_proofs/src/arithmetic.rs `kate_division(a, b)`: quotient of a(X) (n coefficients) by
 * (X - b), n - 1 coefficients to d_q (the remainder a(b) is dropped, as upstream). */
int b200zk_kate_division_dev(const void* d_a, size_t n, const uint64_t b[4], void* d_q, void* stream);

/* out[i] = sum_j coeffs[j] * polys[j][i] over `count` device columns of n elements (coeffs: host,
 * count x 4 limbs): the `acc * y + poly` accumulations of the multi-open provers
 * (halo2_proofs/src/poly/kzg/multiopen/{shplonk,gwc}/prover.rs) in one pass.  d_out may be one
 * of the inputs. */
int b200zk_linear_combination_dev(const void* const* d_polys, const uint64_t* coeffs, uint32_t count, size_t n,
                                  void* d_out, void* stream);

/* ---- wire formats of G1 points (SURVEY.md section 8 f, rank 3) ---------------------------- */
/* halo2curves 0.3.1 src/derive/curve.rs `G1Affine::to_bytes` applied to `count` points: 32 bytes
 * each, x little-endian canonical with the parity of y in bit 7 of byte 31, identity = zeros.
 * `points_xyz`: Jacobian G1 (12 limbs each, e.g. MSM results; normalised on the device);
 * `points_xy`: G1Affine (8 limbs each, e.g. ParamsKZG::g for `ParamsKZG::write`). */
int b200zk_g1_to_bytes(const uint64_t* points_xyz, size_t count, uint8_t* out32);
int b200zk_g1_affine_to_bytes(const uint64_t* points_xy, size_t count, uint8_t* out32);
/* The EVM transcript / calldata encoding (reference solidity_verifier_contract/contract.sol:77-87):
 * 64 bytes per point, x then y big-endian canonical; identity = zeros. */
int b200zk_g1_to_evm_bytes(const uint64_t* points_xyz, size_t count, uint8_t* out64);
/* `G1Affine::from_bytes` for `count` points (the decompression `ParamsKZG::read` performs for
 * every SRS point: one square root each).  Fails, naming the first offending index, when an x
 * is not canonical or not on the curve. */
int b200zk_g1_affine_from_bytes(const uint8_t* in32, size_t count, uint64_t* points_xy);

/* ---- synthetic inputs (benchmark / test support; oracle/bn254.py defines the streams) -- */
int b200zk_gen_scalars_dev(void* d_out, size_t n, uint64_t seed, size_t start);
int b200zk_gen_points_dev(void* d_out, size_t n, uint64_t seed, size_t start);

/* ---- measurement support ------------------------------------------------------------ */
/* Element-wise field arithmetic on the device: out[i] = a[i] op b[i] for n elements of Fr (field = 0) or
 * Fq (field = 1), 4 limbs each.  Operands are taken as given — any representative in [0, 2p), Montgomery
 * form for the arithmetic operations — and the result is the canonical representative.  The known-answer
 * hook for halo2curves 0.3.1 `bn256::{Fr, Fq}` (Mul / Add / Sub / square / invert / neg / double /
 * from_repr / to_repr / pow; [DEP] halo2curves src/bn256/{fr,fq}.rs, reference Cargo.lock:484-486): the
 * kernels of this library inline the same operators.  `b` may be null for the unary operations;
 * B200ZK_FIELD_INV maps 0 to 0; B200ZK_FIELD_POW raises a[i] to the 256-bit integer held in b[i]. */
enum {
    B200ZK_FIELD_MUL = 0, B200ZK_FIELD_ADD = 1, B200ZK_FIELD_SUB = 2, B200ZK_FIELD_SQR = 3, B200ZK_FIELD_INV = 4,
    B200ZK_FIELD_NEG = 5, B200ZK_FIELD_DBL = 6, B200ZK_FIELD_FROM_MONT = 7, B200ZK_FIELD_TO_MONT = 8,
    B200ZK_FIELD_POW = 9, B200ZK_FIELD_OP_LAST = 9
};
int b200zk_field_op(uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);
/* Register-resident Fq multiply chain on every SM; returns field multiplications / s.
 * This is the measured denominator of the MSM integer-pipe roofline. */
int b200zk_modmul_peak(uint32_t iters, double* modmul_per_s_out);
/* Per-stage CUDA-event timing of the MSM pipeline (recorded on the stream the kernels run
 * on).  After a profiled MSM, b200zk_msm_last_stages fills ms_out[0..8] with the stage
 * durations {hist, scan, scatter, sync, accumulate, combine, reduce, reduce-combine, fold}
 * and info_out with {n, window bits c, windows, (bucket, point) pairs, chunk length}. */
int b200zk_msm_profile(int enable);
/* Tuning knobs (benchmarks only): longest chunk of sorted pairs one thread accumulates, longest
 * bucket segment one thread reduces, and a forced window size (0 = cost model). */
int b200zk_msm_tune(uint32_t max_chunk, uint32_t max_seglen, uint32_t force_window_bits);
int b200zk_msm_last_stages(float* ms_out, int capacity, uint64_t info_out[5]);
/* Upload pipeline of b200zk_msm_g1_registered: a single host-side commit of at least `min_n`
 * scalars against a window table is fed in `parts` point ranges that share one bucket set, so
 * the PCIe upload of range p+1 runs under the sort + accumulation of range p (same group
 * element).  The ranges grow geometrically (each 4x the one before: a short first range so that
 * little has to arrive before the GPU starts, the copy being ~4x faster than the computation it
 * feeds).  parts = 1 disables it; parts = 0 restores the defaults: from 2^22 scalars, 3 ranges
 * growing 4x, after which the schedule follows the host link — every piped commit times its
 * copies and itself and the next one uses growth = 0.9 x (commit time / copy time), in 4 ranges
 * below growth 3 (several GPUs copying at once share the host's memory bandwidth).  parts > 0
 * pins the schedule. */
int b200zk_msm_upload_pipeline(uint32_t parts, size_t min_n);
/* The point ranges that pipeline uses for n scalars, `parts` ranges and a growth factor (pure host function, no
 * device needed): begin_out[0 .. *count_out] with begin_out[0] = 0 and begin_out[*count_out] = n; begin_out must
 * hold parts + 1 entries; parts in 1..16.  Fewer ranges than asked for come back when n is small. */
int b200zk_msm_upload_ranges(size_t n, uint32_t parts, double growth, size_t* begin_out, uint32_t* count_out);
/* Transfer pipeline of b200zk_ntt / b200zk_intt on one host buffer of at least 2^min_log_n
 * elements: the first pass runs in `chunks` column ranges, each as soon as its rectangle of the
 * input has arrived, and the last pass likewise, each rectangle of the output leaving while the
 * next range computes (same result).  chunks = 1 disables it.  Defaults: 4 ranges from 2^22. */
int b200zk_ntt_transfer_pipeline(uint32_t chunks, uint32_t min_log_n);
/* Inter-pass twiddles of the NTT: a pass boundary whose 2^(log_n - log_I) distinct twiddles fit
 * 2^direct_twiddle_max_log_n entries reads one table of exactly those powers (one multiplication
 * per element), larger boundaries use two sqrt(n)-entry tables (two multiplications).  Default 20
 * (32 MB, L2-resident).  0 forces the two-table form everywhere (tests pin that path at small
 * sizes).  Cached twiddle tables are dropped. */
int b200zk_ntt_tune(uint32_t direct_twiddle_max_log_n);
/* Number of kernels launched by this library since init (for bench.py gpu_launches). */
uint64_t b200zk_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
