/* oracle/ref_capi.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat C wrapper around the UNMODIFIED reference library (microsoft/SEAL 4.4.3 built from /root/reference by
 * oracle/Makefile into oracle/_ref/libsealref.so).  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can drive the reference's own Evaluator on raw uint64 slabs
 * laid out exactly like seal::Ciphertext::data() ([poly][rns prime][coeff], ciphertext.h:24-37).
 * The product (seal_b200/, include/) never includes, links or loads this.
 */
#ifndef SEALREF_CAPI_H
#define SEALREF_CAPI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sealref_ctx sealref_ctx;

/* scheme: 1 = BFV, 2 = CKKS (seal::scheme_type values, encryptionparams.h).  SEALContext(parms, true, sec none).
 * Keys come from the reference KeyGenerator with a Blake2xb PRNG seeded with {seed,0,..} => deterministic. */
sealref_ctx *sealref_create(int scheme, size_t n, const uint64_t *moduli, size_t k, uint64_t plain_modulus, uint64_t seed);
void sealref_destroy(sealref_ctx *c);
const char *sealref_last_error(void);

/* default parameter helpers (CoeffModulus::Create / BFVDefault / PlainModulus::Batching) */
int sealref_coeff_modulus_create(size_t n, const int *bits, size_t k, uint64_t *out);
int sealref_coeff_modulus_bfv_default(size_t n, uint64_t *out, size_t cap, size_t *k_out);
uint64_t sealref_plain_modulus_batching(size_t n, int bits);

/* tables, for pinning the oracle/product precomputation */
size_t sealref_key_prime_count(const sealref_ctx *c);
int sealref_ntt_root(const sealref_ctx *c, size_t prime_idx, uint64_t *root);
int sealref_ntt_tables(const sealref_ctx *c, size_t prime_idx, uint64_t *root_powers_operand, uint64_t *root_powers_quotient,
                       uint64_t *inv_root_powers_operand, uint64_t *inv_degree);
/* BEHZ auxiliary base at level L: out = [B primes..., m_sk]; returns |Bsk| (0 on error) */
size_t sealref_base_bsk(const sealref_ctx *c, size_t L, uint64_t *out, size_t cap);

/* keys (flattened [digit j][component 2][key prime K][coeff N]) */
int sealref_relin_key(sealref_ctx *c, uint64_t *out);                    /* K-1 digits */
int sealref_galois_key(sealref_ctx *c, uint32_t galois_elt, uint64_t *out); /* generated lazily, cached */
uint32_t sealref_galois_elt_from_step(const sealref_ctx *c, int step);

/* reference Evaluator ops on raw slabs.  L = number of RNS primes the ciphertext carries (selects the level).
 * All return 0 on success, <0 on exception (message via sealref_last_error). */
int sealref_ntt_forward(sealref_ctx *c, size_t L, size_t size, uint64_t *data);   /* Evaluator::transform_to_ntt_inplace */
int sealref_ntt_inverse(sealref_ctx *c, size_t L, size_t size, uint64_t *data);   /* Evaluator::transform_from_ntt_inplace */
int sealref_multiply(sealref_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out3); /* size2 x size2 -> size3 */
int sealref_multiply_sized(sealref_ctx *c, size_t L, size_t size_a, size_t size_b, const uint64_t *a, const uint64_t *b, uint64_t *out); /* general sizes */
int sealref_square(sealref_ctx *c, size_t L, const uint64_t *a, uint64_t *out3);                 /* Evaluator::square_inplace */
int sealref_linear(sealref_ctx *c, int mode, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out); /* 0 add, 1 sub, 2 negate */
int sealref_multiply_plain_ntt(sealref_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *plain, uint64_t *out); /* Evaluator::multiply_plain, both NTT form */
int sealref_secret_key(sealref_ctx *c, uint64_t *out); /* [k][n], NTT form at the key level */
int sealref_decrypt(sealref_ctx *c, size_t L, size_t size, int is_ntt_form, uint64_t correction_factor, const uint64_t *ct, uint64_t *plain); /* Decryptor::decrypt */
int sealref_batch_codec(sealref_ctx *c, int decode, const uint64_t *in, uint64_t *out); /* BatchEncoder::encode (0) / decode (1), n slots */
int sealref_plain_to_ntt(sealref_ctx *c, size_t L, const uint64_t *plain, uint64_t *out); /* transform_to_ntt_inplace(Plaintext, parms_id) */
int sealref_plain_op_coeff(sealref_ctx *c, int mode, size_t L, size_t size, int ct_is_ntt, uint64_t correction_factor, const uint64_t *a, const uint64_t *plain, uint64_t *out); /* 0 multiply_plain, 1 add_plain, 2 sub_plain; coefficient-form plaintext */
int sealref_relinearize(sealref_ctx *c, size_t L, const uint64_t *in3, uint64_t *out2);
int sealref_multiply_relin(sealref_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out2);
int sealref_rescale(sealref_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2);   /* out has L-1 primes (CKKS) */
int sealref_mod_switch(sealref_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2); /* BFV: divide-round; CKKS: drop */
int sealref_apply_galois(sealref_ctx *c, size_t L, const uint64_t *in2, uint32_t galois_elt, uint64_t *out2);
int sealref_rotate(sealref_ctx *c, size_t L, const uint64_t *in2, int step, uint64_t *out2); /* rotate_rows / rotate_vector */

/* semantic round trip helpers (BFV batching): encrypt a vector of n slots, decrypt; for drop-in demos */
/* wire format: Ciphertext::save / load with compr_mode_type::none, EncryptionParameters::parms_id (L == k: key level) */
int sealref_parms_id(sealref_ctx *c, size_t L, uint64_t *out4);
long sealref_ct_save(sealref_ctx *c, size_t L, size_t size, const uint64_t *data, int is_ntt_form, double scale, uint64_t correction_factor, uint8_t *out, size_t capacity);
int sealref_ct_load(sealref_ctx *c, const uint8_t *in, size_t len, uint64_t *data, size_t capacity_words, uint64_t *size, uint64_t *L, int *is_ntt_form, double *scale, uint64_t *correction_factor);
long sealref_kswitch_keys_stream(sealref_ctx *c, uint32_t galois_elt, uint8_t *out, size_t capacity); /* 0: RelinKeys, else GaloisKeys of that element */
long sealref_seeded_ct_stream(sealref_ctx *c, uint8_t *out, size_t capacity);
/* the writers above with an explicit compr_mode (0 = none, 1 = zlib) */
long sealref_ct_save_mode(sealref_ctx *c, size_t L, size_t size, const uint64_t *data, int is_ntt_form, double scale, uint64_t correction_factor, int mode, uint8_t *out, size_t capacity);
long sealref_kswitch_keys_stream_mode(sealref_ctx *c, uint32_t galois_elt, int mode, uint8_t *out, size_t capacity);
long sealref_seeded_ct_stream_mode(sealref_ctx *c, int mode, uint8_t *out, size_t capacity);
/* CKKSEncoder::encode / decode (complex vectors, [count][2] doubles); 1 = the reference threw invalid_argument */
int sealref_ckks_encode(sealref_ctx *c, size_t L, const double *values, size_t count, double scale, uint64_t *out);
int sealref_ckks_decode(sealref_ctx *c, size_t L, const uint64_t *plain, double scale, double *out);
int sealref_public_key(sealref_ctx *c, uint64_t *out);                        /* [2][k][n] */
int sealref_encrypt_zero_asymmetric(sealref_ctx *c, size_t L, uint64_t *out2); /* Encryptor(pk)::encrypt_zero(parms_id of level L): [2][L][n] */
int sealref_encrypt_zero_symmetric(sealref_ctx *c, size_t L, uint64_t *out2); /* Encryptor::encrypt_zero_symmetric(parms_id of level L, ct): [2][L][n] */
int sealref_bfv_encrypt(sealref_ctx *c, const uint64_t *slots, uint64_t *out2);
int sealref_bfv_decrypt(sealref_ctx *c, size_t L, size_t size, const uint64_t *ct, uint64_t *slots, int *noise_budget);

/* CPU baseline: run `op` (0 = multiply+relinearize, 1 = ntt fwd+inv of a size-2 ct, 2 = rotate one step,
 * 3 = rescale, 4 = bfv multiply) on `threads` host threads, each on its own uniform-random ciphertexts with a thread-local
 * memory pool, until every thread has done `reps` ops; returns wall seconds (<0 on error). */
double sealref_time_op(sealref_ctx *c, int op, size_t L, int threads, int reps);

#ifdef __cplusplus
}
#endif
#endif
