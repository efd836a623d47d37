/* oracle/seal_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's (microsoft/SEAL 4.4.3) RNS-polynomial hot path, written with fully
 * reduced `unsigned __int128 %` arithmetic (no Shoup/Barrett/lazy ranges) so that it is an independent statement of
 * WHAT the reference computes.  Every function cites the reference file:line it follows (paths relative to
 * /root/reference/native/src/seal/).
 *
 * PINNING: tests/test_oracle*.py check this file against (i) the reference's in-source known-answer tests
 * (tests/seal/util/ntt.cpp:53-133, tests/seal/util/galois.cpp:28-120, tests/seal/util/rns.cpp:347-438 BaseConverter,
 * :460-853 FastBConvMTilde / MontgomeryReduction / FastFloor / FastBConvSK, :904-1011 DivideAndRoundQLastInplace),
 * (ii) golden vectors generated from the real
 * reference in the build container (tests/golden/, generator tests/golden/make_golden.py) and (iii) live outputs of
 * oracle/_ref/libsealref.so (the reference compiled from its own sources) wherever that library is present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the product never does.
 */
#ifndef SEAL_ORACLE_H
#define SEAL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_BFV 1
#define ORC_CKKS 2
#define ORC_BGV 3 /* NTT-form ciphertexts like CKKS (orc_ckks_multiply is bgv_multiply's tensor, evaluator.cpp:710-841) */
#define ORC_MAX_PRIMES 64

typedef struct orc_ctx orc_ctx;

/* number theory (util/numth.cpp) */
int orc_is_prime(uint64_t v);
int orc_get_primes(uint64_t factor, int bit_size, size_t count, uint64_t *out);           /* numth.cpp:278-311 */
int orc_minimal_primitive_root(uint64_t degree, uint64_t q, uint64_t *root);              /* numth.cpp:340-412 */
int orc_coeff_modulus_create(size_t n, const int *bits, size_t k, uint64_t *out);         /* modulus.cpp:144-184 */

/* context: moduli = the key-level coeff_modulus (k primes, last = special prime); t = plain modulus (BFV, BGV) */
orc_ctx *orc_create(int scheme, size_t n, const uint64_t *moduli, size_t k, uint64_t t);
void orc_destroy(orc_ctx *c);
/* tables of prime i, ntt.cpp:241-300 */
int orc_ntt_tables(const orc_ctx *c, size_t i, uint64_t *root, uint64_t *root_powers, uint64_t *inv_root_powers, uint64_t *inv_n);
/* BEHZ auxiliary base at level L: [B..., m_sk]; returns |Bsk| (rns.cpp:598-641) */
size_t orc_base_bsk(const orc_ctx *c, size_t L, uint64_t *out);

/* RNS base conversion building blocks with explicit bases (so the reference's tiny-base KATs can be replayed) */
void orc_fastbconv_array(const uint64_t *ibase, size_t ni, const uint64_t *obase, size_t no, const uint64_t *in, size_t n, uint64_t *out); /* rns.cpp:418-463 */
size_t orc_behz_base(size_t n, const uint64_t *q, size_t L, uint64_t t, uint64_t *bsk_out);                      /* rns.cpp:598-641 */
int orc_behz_fastbconv_m_tilde(size_t n, const uint64_t *q, size_t L, uint64_t t, const uint64_t *in, uint64_t *out); /* rns.cpp:1086-1131 */
int orc_behz_sm_mrq(size_t n, const uint64_t *q, size_t L, uint64_t t, const uint64_t *in, uint64_t *out);      /* rns.cpp:979-1039 */
int orc_behz_fast_floor(size_t n, const uint64_t *q, size_t L, uint64_t t, const uint64_t *in, uint64_t *out);  /* rns.cpp:1041-1084 */
int orc_behz_fastbconv_sk(size_t n, const uint64_t *q, size_t L, uint64_t t, const uint64_t *in, uint64_t *out); /* rns.cpp:903-977 */
void orc_divide_and_round_q_last(const uint64_t *q, size_t L, size_t n, uint64_t *data);                        /* rns.cpp:789-828 */

/* single-row transforms with explicit modulus index (ntt.cpp:394-475 / dwthandler.h:94-356); canonical outputs */
void orc_ntt_row(const orc_ctx *c, size_t prime_idx, uint64_t *row);
void orc_intt_row(const orc_ctx *c, size_t prime_idx, uint64_t *row);

/* slab ops; slabs are [size][L][n] like seal::Ciphertext::data() */
void orc_ntt_forward(const orc_ctx *c, size_t L, size_t size, uint64_t *data);            /* evaluator.cpp:2289-2335 */
void orc_ntt_inverse(const orc_ctx *c, size_t L, size_t size, uint64_t *data);            /* evaluator.cpp:2337-2382 */
void orc_ckks_multiply(const orc_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out3); /* evaluator.cpp:569-708 */
/* add (mode 0), sub (1), negate (2) on [size][L][n]; evaluator.cpp:130-350 */
void orc_linear(const orc_ctx *c, int mode, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* Evaluator::multiply_plain, ciphertext and plaintext in NTT form (evaluator.cpp:2157-2195) */
void orc_multiply_plain_ntt(const orc_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *plain, uint64_t *out);
/* BatchEncoder::encode (decode = 0) / decode (1) on n matrix slots (batchencoder.cpp:54-130, :229-275); -1: t does not support batching */
int orc_batch_codec(const orc_ctx *c, int decode, const uint64_t *in, uint64_t *out);
/* coefficient-form plaintexts (n words < t): transform_to_ntt(Plaintext) :2197-2287, multiply_plain :2021-2155 / :1999-2004,
 * add_plain / sub_plain util/scalingvariant.cpp:70-160 (BFV) and evaluator.cpp:1838-1849 (BGV) */
void orc_plain_to_ntt(const orc_ctx *c, size_t L, const uint64_t *plain, uint64_t *out);
void orc_multiply_plain_coeff(const orc_ctx *c, size_t L, size_t size, int ct_is_ntt, const uint64_t *a, const uint64_t *plain, uint64_t *out);
void orc_add_plain_coeff(const orc_ctx *c, size_t L, size_t size, int subtract, uint64_t correction_factor, const uint64_t *a, const uint64_t *plain, uint64_t *out);
int orc_bfv_multiply(const orc_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out3);   /* evaluator.cpp:395-567 */
/* general ciphertext sizes s1 x s2 -> s1+s2-1 (evaluator.cpp:664-700, :796-833; BFV :453-560) */
void orc_ckks_multiply_sized(const orc_ctx *c, size_t L, size_t s1, size_t s2, const uint64_t *a, const uint64_t *b, uint64_t *out);
int orc_bfv_multiply_sized(const orc_ctx *c, size_t L, size_t s1, size_t s2, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* ct (size 2, updated in place) += key-switch of target ([L][n]); key = [L digits][2][k][n]; evaluator.cpp:2561-2867 */
void orc_switch_key(const orc_ctx *c, size_t L, uint64_t *ct2, const uint64_t *target, const uint64_t *key);
void orc_relinearize(const orc_ctx *c, size_t L, const uint64_t *in3, const uint64_t *key, uint64_t *out2); /* evaluator.cpp:1144-1199 */
void orc_rescale(const orc_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2);        /* CKKS; rns.cpp:830-901, evaluator.cpp:1201-1294 */
void orc_bfv_mod_switch(const orc_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2); /* BFV; rns.cpp:789-828 */
void orc_bgv_mod_switch(const orc_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2); /* BGV; rns.cpp:1193-1236 */
void orc_apply_galois(const orc_ctx *c, size_t L, const uint64_t *in2, uint32_t galois_elt, const uint64_t *key, uint64_t *out2); /* evaluator.cpp:2384-2502 */
uint32_t orc_galois_elt_from_step(size_t n, int step);                                    /* galois.cpp:53-95 */
/* bare permutations (galois.cpp:148-218) on one row */
void orc_galois_coeff_row(size_t n, uint64_t q, uint32_t galois_elt, const uint64_t *in, uint64_t *out);
void orc_galois_ntt_row(size_t n, uint32_t galois_elt, const uint64_t *in, uint64_t *out);

/* decryption (decryptor.cpp); sk = the secret key in NTT form at the key level, [k][n] */
void orc_decrypt_phase(const orc_ctx *c, size_t L, size_t size, int ct_is_ntt, const uint64_t *ct, const uint64_t *sk, uint64_t *out); /* :312-384; CKKS decrypt = this */
int orc_bfv_decrypt(const orc_ctx *c, size_t L, size_t size, const uint64_t *ct, const uint64_t *sk, uint64_t *plain);                 /* :111-135, rns.cpp:1133-1191 */
int orc_bgv_decrypt(const orc_ctx *c, size_t L, size_t size, uint64_t correction_factor, const uint64_t *ct, const uint64_t *sk, uint64_t *plain); /* :159-197, rns.cpp:466-539 */

/* seed-compressed ciphertexts: Ciphertext::expand_seed (ciphertext.cpp:118-150) = sample_poly_uniform (util/rlwe.cpp:104-132) on a
 * Blake2xbPRNG (randomgen.cpp:204-214: buffer b = BLAKE2Xb(4096 bytes, in = counter b, key = the 64-byte seed), util/blake2xb.c).
 * seed = prng_seed_type (8 words); out = the polynomial [L][n] the seed expands into at the level with L primes */
void orc_blake2xb_stream(const uint64_t seed[8], size_t words, uint64_t *out); /* the first `words` 64-bit words of the PRNG's output */
void orc_expand_seed(const orc_ctx *c, size_t L, const uint64_t seed[8], uint64_t *out);
/* Encryptor::encrypt_zero (public key; encryptor.cpp:88-174, util/rlwe.cpp:184-276) with the PRNG seeded by `seed`; pk = [2][k][n]
 * NTT form; out = [2][L][n] at the level with L primes (sampled one level above and divided down, L == k: sampled directly) */
void orc_encrypt_zero_asymmetric(const orc_ctx *c, size_t L, const uint64_t *pk, const uint64_t seed[8], uint64_t *out2);
/* CKKSEncoder::encode / decode of complex vectors (ckks.h:455-807): values = [count][2] doubles, plaintext = [L][n] NTT form;
 * return 0, or 1 for the reference's invalid_argument cases (values too large, scale out of bounds, non-finite input) */
int orc_ckks_encode(const orc_ctx *c, size_t L, const double *values, size_t count, double scale, uint64_t *out);
int orc_ckks_decode(const orc_ctx *c, size_t L, const uint64_t *plain, double scale, double *out);
/* encrypt_zero_symmetric (util/rlwe.cpp:264-408) with the bootstrap PRNG seeded by `seed`; sk = [k][n] NTT form; out = [2][L][n] at
 * the first data level; save_seed selects which BFV variant (the Serializable one or the plain one) */
void orc_encrypt_zero_symmetric(const orc_ctx *c, size_t L, const uint64_t *sk, const uint64_t seed[8], int save_seed, uint64_t *out2);

#ifdef __cplusplus
}
#endif
#endif
