/* b200zk_witness — native, multithreaded witness values for the reference's RSA modular
 * exponentiation (SURVEY.md section 8 f, rank 4).  Host-only (no CUDA): it sits beside the hot
 * path, where `create_proof`'s caller synthesises the witness.
 *
 * Replaces the BigUint arithmetic inside
 *   reference src/big_uint/chip.rs:454-490  `BigUintConfig::pow_mod_fixed_exp(a, e, n)`
 *   reference src/big_uint/chip.rs:355-413  `mul_mod` / `square_mod`
 *   reference src/big_uint/chip.rs:513-608  `is_equal_muled` (carries of the running sum)
 * reached from `RSAConfig::modpow_public_key` (reference src/chip.rs:81-96) in
 * `verify_pkcs1v15_signature` (src/chip.rs:110-236).  Integers are little-endian 64-bit limbs
 * (the circuit's limb_bits = 64, reference src/lib.rs:268).
 *
 * One mul_mod record, in 64-bit words (L = num_limbs, M = 2 L - 1):
 *   a[L] b[L] q[L] r[L] ab[M][3] qn[M][3] carry[M][2] c[M]
 *     q, r       quotient and remainder of a * b by n                      (chip.rs:372-382)
 *     ab, qn     carry-less limb products sum_{i+j=k} a_i b_j and q_i n_j  (chip.rs:383-384)
 *     carry, c   upper bits and low 64 bits of ab_k - (qn_k + r_k) + carry_k + muled_limb_max
 *                                                                          (chip.rs:551-570, :752-756)
 * One exponentiation = b200zk_witness_pow_steps(e) records in the chip's order — per bit of e from
 * the lowest: square_mod(cur_sq), then mul_mod(acc, cur_sq) when the bit is set — followed by the
 * result a^e mod n [L]. */
#ifndef B200ZK_WITNESS_H
#define B200ZK_WITNESS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

size_t b200zk_witness_mul_mod_words(uint32_t num_limbs);
uint32_t b200zk_witness_pow_steps(uint64_t e);
size_t b200zk_witness_pow_words(uint64_t e, uint32_t num_limbs);

/* `count` independent exponentiations (base[i], modulus[i], each num_limbs words, base < modulus)
 * on `threads` host threads (<= 0: all); out holds count * b200zk_witness_pow_words(e, num_limbs)
 * words.  Returns 0, 1 for bad arguments, 2 when some base >= modulus (the chip asserts x < n,
 * reference src/chip.rs:88), 3 when an identity the circuit checks does not hold (cannot happen). */
int b200zk_witness_pow_mod_fixed_exp(const uint64_t* base, const uint64_t* modulus, uint64_t e, uint32_t num_limbs,
                                     size_t count, int threads, uint64_t* out);

/* `count` canonical integers of `width` (1..4) words, each below the BN254 scalar modulus, to
 * `bn256::Fr` Montgomery limbs (4 words each): the form advice cells take on the device. */
int b200zk_witness_words_to_fr(const uint64_t* words, uint32_t width, size_t count, int threads, uint64_t* out_fr);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_WITNESS_H */
