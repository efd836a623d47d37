// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (see ref_capi.h).
// Drives the unmodified reference (seal::Evaluator etc., compiled from /root/reference by oracle/Makefile) on raw
// uint64 slabs.  This file contains no arithmetic of its own: every number it returns is computed by the reference.
#include "ref_capi.h"
#include "seal/seal.h"
#include <chrono>
#include <complex>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

using namespace seal;

static thread_local std::string g_err;

struct sealref_ctx
{
    std::unique_ptr<SEALContext> context;
    std::unique_ptr<KeyGenerator> keygen;
    std::unique_ptr<Evaluator> evaluator;
    std::unique_ptr<RelinKeys> relin;
    std::map<uint32_t, GaloisKeys> galois;
    scheme_type scheme;
    size_t n = 0;
    size_t k = 0;
};

#define REF_TRY try {
#define REF_CATCH(ret)                  \
    }                                   \
    catch (const std::exception &e)     \
    {                                   \
        g_err = e.what();               \
        return ret;                     \
    }

extern "C" const char *sealref_last_error(void)
{
    return g_err.c_str();
}

extern "C" sealref_ctx *sealref_create(
    int scheme, size_t n, const uint64_t *moduli, size_t k, uint64_t plain_modulus, uint64_t seed)
{
    REF_TRY
    auto c = std::make_unique<sealref_ctx>();
    c->scheme = static_cast<scheme_type>(scheme);
    EncryptionParameters parms(c->scheme);
    parms.set_poly_modulus_degree(n);
    std::vector<Modulus> mods;
    for (size_t i = 0; i < k; i++)
        mods.emplace_back(moduli[i]);
    parms.set_coeff_modulus(mods);
    if (c->scheme == scheme_type::bfv || c->scheme == scheme_type::bgv)
        parms.set_plain_modulus(plain_modulus);
    prng_seed_type s{};
    s[0] = seed;
    parms.set_random_generator(std::make_shared<Blake2xbPRNGFactory>(s));
    c->context = std::make_unique<SEALContext>(parms, true, sec_level_type::none);
    if (!c->context->parameters_set())
    {
        g_err = std::string("invalid parameters: ") + c->context->parameter_error_message();
        return nullptr;
    }
    c->keygen = std::make_unique<KeyGenerator>(*c->context);
    c->evaluator = std::make_unique<Evaluator>(*c->context);
    c->n = n;
    c->k = k;
    return c.release();
    REF_CATCH(nullptr)
}

extern "C" void sealref_destroy(sealref_ctx *c)
{
    delete c;
}

extern "C" int sealref_coeff_modulus_create(size_t n, const int *bits, size_t k, uint64_t *out)
{
    REF_TRY
    auto v = CoeffModulus::Create(n, std::vector<int>(bits, bits + k));
    for (size_t i = 0; i < k; i++)
        out[i] = v[i].value();
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_coeff_modulus_bfv_default(size_t n, uint64_t *out, size_t cap, size_t *k_out)
{
    REF_TRY
    auto v = CoeffModulus::BFVDefault(n);
    if (v.size() > cap)
        return -2;
    for (size_t i = 0; i < v.size(); i++)
        out[i] = v[i].value();
    *k_out = v.size();
    return 0;
    REF_CATCH(-1)
}

extern "C" uint64_t sealref_plain_modulus_batching(size_t n, int bits)
{
    REF_TRY
    return PlainModulus::Batching(n, bits).value();
    REF_CATCH(0)
}

extern "C" size_t sealref_key_prime_count(const sealref_ctx *c)
{
    return c->k;
}

extern "C" int sealref_ntt_root(const sealref_ctx *c, size_t prime_idx, uint64_t *root)
{
    REF_TRY
    *root = c->context->key_context_data()->small_ntt_tables()[prime_idx].get_root();
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_ntt_tables(
    const sealref_ctx *c, size_t prime_idx, uint64_t *rp_op, uint64_t *rp_quo, uint64_t *irp_op, uint64_t *inv_degree)
{
    REF_TRY
    auto &t = c->context->key_context_data()->small_ntt_tables()[prime_idx];
    for (size_t i = 0; i < c->n; i++)
    {
        rp_op[i] = t.get_from_root_powers()[i].operand;
        rp_quo[i] = t.get_from_root_powers()[i].quotient;
        irp_op[i] = t.get_from_inv_root_powers()[i].operand;
    }
    *inv_degree = t.inv_degree_modulo().operand;
    return 0;
    REF_CATCH(-1)
}

static std::shared_ptr<const SEALContext::ContextData> level(const sealref_ctx *c, size_t L)
{
    if (L == c->k && c->k > 1)
        return c->context->key_context_data();
    auto cd = c->context->first_context_data();
    while (cd && cd->parms().coeff_modulus().size() != L)
        cd = cd->next_context_data();
    if (!cd)
        throw std::invalid_argument("no level with that many primes");
    return cd;
}

extern "C" size_t sealref_base_bsk(const sealref_ctx *c, size_t L, uint64_t *out, size_t cap)
{
    REF_TRY
    auto cd = level(c, L);
    auto bsk = cd->rns_tool()->base_Bsk();
    if (bsk->size() > cap)
        return 0;
    for (size_t i = 0; i < bsk->size(); i++)
        out[i] = (*bsk)[i].value();
    return bsk->size();
    REF_CATCH(0)
}

static void flatten_key(const sealref_ctx *c, const std::vector<PublicKey> &kv, uint64_t *out)
{
    size_t row = c->k * c->n;
    for (size_t j = 0; j < kv.size(); j++)
    {
        const Ciphertext &ct = kv[j].data();
        if (ct.size() != 2 || ct.coeff_modulus_size() != c->k)
            throw std::logic_error("unexpected key shape");
        std::memcpy(out + j * 2 * row, ct.data(), 2 * row * sizeof(uint64_t));
    }
}

extern "C" int sealref_relin_key(sealref_ctx *c, uint64_t *out)
{
    REF_TRY
    if (!c->relin)
    {
        c->relin = std::make_unique<RelinKeys>();
        c->keygen->create_relin_keys(*c->relin);
    }
    flatten_key(c, c->relin->key(2), out);
    return 0;
    REF_CATCH(-1)
}

static const GaloisKeys &galois_for(sealref_ctx *c, uint32_t elt)
{
    auto it = c->galois.find(elt);
    if (it == c->galois.end())
    {
        GaloisKeys g;
        c->keygen->create_galois_keys(std::vector<uint32_t>{ elt }, g);
        it = c->galois.emplace(elt, std::move(g)).first;
    }
    return it->second;
}

extern "C" int sealref_galois_key(sealref_ctx *c, uint32_t galois_elt, uint64_t *out)
{
    REF_TRY
    flatten_key(c, galois_for(c, galois_elt).key(galois_elt), out);
    return 0;
    REF_CATCH(-1)
}

extern "C" uint32_t sealref_galois_elt_from_step(const sealref_ctx *c, int step)
{
    REF_TRY
    return c->context->key_context_data()->galois_tool()->get_elt_from_step(step);
    REF_CATCH(0)
}

static double default_scale(const SEALContext::ContextData &cd)
{
    // any in-bounds value; the scale is metadata and does not touch the residues
    (void)cd;
    return 1024.0;
}

static Ciphertext make_ct(
    const sealref_ctx *c, size_t L, size_t size, const uint64_t *data, MemoryPoolHandle pool = MemoryManager::GetPool())
{
    auto cd = level(c, L);
    Ciphertext ct(pool);
    ct.resize(*c->context, cd->parms_id(), size);
    ct.is_ntt_form() = (c->scheme != scheme_type::bfv);
    if (c->scheme == scheme_type::ckks)
        ct.scale() = default_scale(*cd);
    if (data)
        std::memcpy(ct.data(), data, size * L * c->n * sizeof(uint64_t));
    return ct;
}

static void store_ct(const sealref_ctx *c, const Ciphertext &ct, uint64_t *out)
{
    std::memcpy(out, ct.data(), ct.size() * ct.coeff_modulus_size() * c->n * sizeof(uint64_t));
}

extern "C" int sealref_ntt_forward(sealref_ctx *c, size_t L, size_t size, uint64_t *data)
{
    REF_TRY
    Ciphertext ct = make_ct(c, L, size, data);
    ct.is_ntt_form() = false;
    c->evaluator->transform_to_ntt_inplace(ct);
    store_ct(c, ct, data);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_ntt_inverse(sealref_ctx *c, size_t L, size_t size, uint64_t *data)
{
    REF_TRY
    Ciphertext ct = make_ct(c, L, size, data);
    ct.is_ntt_form() = true;
    c->evaluator->transform_from_ntt_inplace(ct);
    store_ct(c, ct, data);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_multiply(sealref_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out3)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, a), y = make_ct(c, L, 2, b);
    c->evaluator->multiply_inplace(x, y);
    store_ct(c, x, out3);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_multiply_sized(
    sealref_ctx *c, size_t L, size_t size_a, size_t size_b, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, size_a, a), y = make_ct(c, L, size_b, b);
    c->evaluator->multiply_inplace(x, y);
    store_ct(c, x, out);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_square(sealref_ctx *c, size_t L, const uint64_t *a, uint64_t *out3)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, a);
    c->evaluator->square_inplace(x);
    store_ct(c, x, out3);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_linear(sealref_ctx *c, int mode, size_t L, size_t size, const uint64_t *a, const uint64_t *b, uint64_t *out)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, size, a);
    if (mode == 2)
        c->evaluator->negate_inplace(x);
    else
    {
        Ciphertext y = make_ct(c, L, size, b);
        if (mode == 0)
            c->evaluator->add_inplace(x, y);
        else
            c->evaluator->sub_inplace(x, y);
    }
    store_ct(c, x, out);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_multiply_plain_ntt(sealref_ctx *c, size_t L, size_t size, const uint64_t *a, const uint64_t *plain, uint64_t *out)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, size, a);
    x.is_ntt_form() = true; // BFV callers pass an already transformed ciphertext (evaluator.cpp:1991-1994)
    auto cd = level(c, L);
    Plaintext p;
    p.resize(L * c->n);
    std::memcpy(p.data(), plain, L * c->n * sizeof(uint64_t));
    p.parms_id() = cd->parms_id(); // NTT-form plaintext at the ciphertext's level
    p.scale() = c->scheme == scheme_type::ckks ? default_scale(*cd) : 1.0;
    c->evaluator->multiply_plain_inplace(x, p);
    store_ct(c, x, out);
    return 0;
    REF_CATCH(-1)
}

// the secret key (NTT form, key level) and Decryptor::decrypt of an arbitrary ciphertext; plain gets n words (BFV / BGV,
// zero-padded) or L*n words (CKKS, NTT form)
extern "C" int sealref_secret_key(sealref_ctx *c, uint64_t *out)
{
    REF_TRY
    std::memcpy(out, c->keygen->secret_key().data().data(), c->k * c->n * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_decrypt(
    sealref_ctx *c, size_t L, size_t size, int is_ntt_form, uint64_t correction_factor, const uint64_t *ct, uint64_t *plain)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, size, ct);
    x.is_ntt_form() = is_ntt_form != 0;
    x.correction_factor() = correction_factor;
    Decryptor dec(*c->context, c->keygen->secret_key());
    Plaintext p;
    dec.decrypt(x, p);
    size_t words = c->scheme == scheme_type::ckks ? L * c->n : c->n;
    std::memset(plain, 0, words * sizeof(uint64_t));
    std::memcpy(plain, p.data(), std::min(words, p.coeff_count()) * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

// BatchEncoder::encode / decode on n matrix slots
extern "C" int sealref_batch_codec(sealref_ctx *c, int decode, const uint64_t *in, uint64_t *out)
{
    REF_TRY
    BatchEncoder enc(*c->context);
    if (decode)
    {
        Plaintext p(c->n);
        std::memcpy(p.data(), in, c->n * sizeof(uint64_t));
        std::vector<uint64_t> v;
        enc.decode(p, v);
        std::memcpy(out, v.data(), c->n * sizeof(uint64_t));
    }
    else
    {
        Plaintext p;
        enc.encode(std::vector<uint64_t>(in, in + c->n), p);
        std::memset(out, 0, c->n * sizeof(uint64_t));
        std::memcpy(out, p.data(), p.coeff_count() * sizeof(uint64_t));
    }
    return 0;
    REF_CATCH(-1)
}

// coefficient-form plaintext of n words (< plain_modulus)
static Plaintext make_plain(const sealref_ctx *c, const uint64_t *words)
{
    Plaintext p(c->n);
    std::memcpy(p.data(), words, c->n * sizeof(uint64_t));
    return p;
}

extern "C" int sealref_plain_to_ntt(sealref_ctx *c, size_t L, const uint64_t *plain, uint64_t *out)
{
    REF_TRY
    Plaintext p = make_plain(c, plain);
    c->evaluator->transform_to_ntt_inplace(p, level(c, L)->parms_id());
    std::memcpy(out, p.data(), L * c->n * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

// Evaluator::multiply_plain / add_plain / sub_plain (mode 0 / 1 / 2) with a coefficient-form plaintext
extern "C" int sealref_plain_op_coeff(
    sealref_ctx *c, int mode, size_t L, size_t size, int ct_is_ntt, uint64_t correction_factor, const uint64_t *a,
    const uint64_t *plain, uint64_t *out)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, size, a);
    x.is_ntt_form() = ct_is_ntt != 0;
    x.correction_factor() = correction_factor;
    Plaintext p = make_plain(c, plain);
    if (mode == 0)
        c->evaluator->multiply_plain_inplace(x, p);
    else if (mode == 1)
        c->evaluator->add_plain_inplace(x, p);
    else
        c->evaluator->sub_plain_inplace(x, p);
    store_ct(c, x, out);
    return 0;
    REF_CATCH(-1)
}

static const RelinKeys &relin_keys(sealref_ctx *c)
{
    if (!c->relin)
    {
        c->relin = std::make_unique<RelinKeys>();
        c->keygen->create_relin_keys(*c->relin);
    }
    return *c->relin;
}

extern "C" int sealref_relinearize(sealref_ctx *c, size_t L, const uint64_t *in3, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 3, in3);
    c->evaluator->relinearize_inplace(x, relin_keys(c));
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_multiply_relin(sealref_ctx *c, size_t L, const uint64_t *a, const uint64_t *b, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, a), y = make_ct(c, L, 2, b);
    c->evaluator->multiply_inplace(x, y);
    c->evaluator->relinearize_inplace(x, relin_keys(c));
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_rescale(sealref_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, in2);
    // scale must stay in bounds after division by q_last: use q_last * 1024
    x.scale() = static_cast<double>(level(c, L)->parms().coeff_modulus().back().value()) * 1024.0;
    c->evaluator->rescale_to_next_inplace(x);
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_mod_switch(sealref_ctx *c, size_t L, const uint64_t *in2, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, in2);
    c->evaluator->mod_switch_to_next_inplace(x);
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_apply_galois(sealref_ctx *c, size_t L, const uint64_t *in2, uint32_t galois_elt, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, in2);
    c->evaluator->apply_galois_inplace(x, galois_elt, galois_for(c, galois_elt));
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_rotate(sealref_ctx *c, size_t L, const uint64_t *in2, int step, uint64_t *out2)
{
    REF_TRY
    Ciphertext x = make_ct(c, L, 2, in2);
    uint32_t elt = c->context->key_context_data()->galois_tool()->get_elt_from_step(step);
    const GaloisKeys &g = galois_for(c, elt);
    if (c->scheme == scheme_type::ckks)
        c->evaluator->rotate_vector_inplace(x, step, g);
    else
        c->evaluator->rotate_rows_inplace(x, step, g);
    store_ct(c, x, out2);
    return 0;
    REF_CATCH(-1)
}

// ---- wire format (Ciphertext::save / load, compr_mode_type::none) and parms_id, for the rank-3 parity tests ----
extern "C" int sealref_parms_id(sealref_ctx *c, size_t L, uint64_t *out4)
{
    REF_TRY
    auto id = L == c->k ? c->context->key_parms_id() : level(c, L)->parms_id();
    for (int i = 0; i < 4; i++)
        out4[i] = id[i];
    return 0;
    REF_CATCH(-1)
}

extern "C" long sealref_ct_save(
    sealref_ctx *c, size_t L, size_t size, const uint64_t *data, int is_ntt_form, double scale, uint64_t correction_factor,
    uint8_t *out, size_t capacity)
{
    REF_TRY
    Ciphertext ct = make_ct(c, L, size, data);
    ct.is_ntt_form() = is_ntt_form != 0;
    ct.scale() = scale;
    ct.correction_factor() = correction_factor;
    return static_cast<long>(ct.save(reinterpret_cast<seal_byte *>(out), capacity, compr_mode_type::none));
    REF_CATCH(-1)
}

// the same three writers with an explicit compression mode (0 none, 1 zlib; serialization.h:33-47)
static compr_mode_type mode_of(int mode)
{
    if (mode == 0)
        return compr_mode_type::none;
#ifdef SEAL_USE_ZLIB
    if (mode == 1)
        return compr_mode_type::zlib;
#endif
    throw std::invalid_argument("unsupported compression mode");
}
extern "C" long sealref_ct_save_mode(
    sealref_ctx *c, size_t L, size_t size, const uint64_t *data, int is_ntt_form, double scale, uint64_t correction_factor, int mode,
    uint8_t *out, size_t capacity)
{
    REF_TRY
    Ciphertext ct = make_ct(c, L, size, data);
    ct.is_ntt_form() = is_ntt_form != 0;
    ct.scale() = scale;
    ct.correction_factor() = correction_factor;
    return static_cast<long>(ct.save(reinterpret_cast<seal_byte *>(out), capacity, mode_of(mode)));
    REF_CATCH(-1)
}
extern "C" long sealref_kswitch_keys_stream_mode(sealref_ctx *c, uint32_t galois_elt, int mode, uint8_t *out, size_t capacity)
{
    REF_TRY
    if (galois_elt == 0)
        return static_cast<long>(relin_keys(c).save(reinterpret_cast<seal_byte *>(out), capacity, mode_of(mode)));
    return static_cast<long>(galois_for(c, galois_elt).save(reinterpret_cast<seal_byte *>(out), capacity, mode_of(mode)));
    REF_CATCH(-1)
}
extern "C" long sealref_seeded_ct_stream_mode(sealref_ctx *c, int mode, uint8_t *out, size_t capacity)
{
    REF_TRY
    Encryptor encryptor(*c->context, c->keygen->secret_key());
    auto ser = encryptor.encrypt_zero_symmetric();
    return static_cast<long>(ser.save(reinterpret_cast<seal_byte *>(out), capacity, mode_of(mode)));
    REF_CATCH(-1)
}

extern "C" int sealref_ct_load(
    sealref_ctx *c, const uint8_t *in, size_t len, uint64_t *data, size_t capacity_words, uint64_t *size, uint64_t *L,
    int *is_ntt_form, double *scale, uint64_t *correction_factor)
{
    REF_TRY
    Ciphertext ct;
    ct.load(*c->context, reinterpret_cast<const seal_byte *>(in), len);
    size_t words = ct.size() * ct.coeff_modulus_size() * ct.poly_modulus_degree();
    if (words > capacity_words)
        throw std::invalid_argument("capacity");
    std::memcpy(data, ct.data(), words * sizeof(uint64_t));
    *size = ct.size(), *L = ct.coeff_modulus_size(), *is_ntt_form = ct.is_ntt_form(), *scale = ct.scale();
    *correction_factor = ct.correction_factor();
    return 0;
    REF_CATCH(-1)
}

// RelinKeys / GaloisKeys::save(compr_mode_type::none); galois_elt == 0 selects the relinearization keys
extern "C" long sealref_kswitch_keys_stream(sealref_ctx *c, uint32_t galois_elt, uint8_t *out, size_t capacity)
{
    REF_TRY
    if (galois_elt == 0)
        return static_cast<long>(relin_keys(c).save(reinterpret_cast<seal_byte *>(out), capacity, compr_mode_type::none));
    return static_cast<long>(galois_for(c, galois_elt).save(reinterpret_cast<seal_byte *>(out), capacity, compr_mode_type::none));
    REF_CATCH(-1)
}

// a seed-compressed ciphertext as Encryptor::encrypt_zero_symmetric(...).save() writes it (c_1 replaced by its PRNG seed)
extern "C" long sealref_seeded_ct_stream(sealref_ctx *c, uint8_t *out, size_t capacity)
{
    REF_TRY
    Encryptor encryptor(*c->context, c->keygen->secret_key());
    auto ser = encryptor.encrypt_zero_symmetric();
    return static_cast<long>(ser.save(reinterpret_cast<seal_byte *>(out), capacity, compr_mode_type::none));
    REF_CATCH(-1)
}

// the public key (KeyGenerator::create_public_key: [2][k][n], NTT form at the key level) and Encryptor::encrypt_zero(parms_id, ct)
// with it; L == k: the key level itself.  Deterministic: every PRNG the seeded factory creates starts from {seed, 0, ..., 0}
extern "C" int sealref_public_key(sealref_ctx *c, uint64_t *out)
{
    REF_TRY
    PublicKey pk;
    c->keygen->create_public_key(pk);
    std::memcpy(out, pk.data().data(), 2 * c->k * c->n * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_encrypt_zero_asymmetric(sealref_ctx *c, size_t L, uint64_t *out2)
{
    REF_TRY
    PublicKey pk;
    c->keygen->create_public_key(pk);
    Encryptor encryptor(*c->context, pk);
    Ciphertext ct;
    encryptor.encrypt_zero(L == c->k ? c->context->key_parms_id() : level(c, L)->parms_id(), ct);
    std::memcpy(out2, ct.data(), 2 * L * c->n * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

// CKKSEncoder::encode(vector<complex<double>>, parms_id, scale, plain) / decode; values = [count][2] doubles.
// Returns 1 when the reference throws invalid_argument (values too large, scale out of bounds, non-finite input).
extern "C" int sealref_ckks_encode(sealref_ctx *c, size_t L, const double *values, size_t count, double scale, uint64_t *out)
{
    try
    {
        CKKSEncoder encoder(*c->context);
        std::vector<std::complex<double>> v(count);
        for (size_t i = 0; i < count; i++)
            v[i] = { values[2 * i], values[2 * i + 1] };
        Plaintext p;
        encoder.encode(v, level(c, L)->parms_id(), scale, p);
        std::memcpy(out, p.data(), L * c->n * sizeof(uint64_t));
        return 0;
    }
    catch (const std::invalid_argument &e)
    {
        g_err = e.what();
        return 1;
    }
    catch (const std::exception &e)
    {
        g_err = e.what();
        return -1;
    }
}

extern "C" int sealref_ckks_decode(sealref_ctx *c, size_t L, const uint64_t *plain, double scale, double *out)
{
    try
    {
        CKKSEncoder encoder(*c->context);
        auto cd = level(c, L);
        Plaintext p;
        p.resize(L * c->n);
        std::memcpy(p.data(), plain, L * c->n * sizeof(uint64_t));
        p.parms_id() = cd->parms_id();
        p.scale() = scale;
        std::vector<std::complex<double>> v;
        encoder.decode(p, v);
        for (size_t i = 0; i < v.size(); i++)
            out[2 * i] = v[i].real(), out[2 * i + 1] = v[i].imag();
        return 0;
    }
    catch (const std::invalid_argument &e)
    {
        g_err = e.what();
        return 1;
    }
    catch (const std::exception &e)
    {
        g_err = e.what();
        return -1;
    }
}

// Encryptor::encrypt_zero_symmetric(destination): the plain (not seed-compressed) variant.  Deterministic here: the context's random
// generator factory is seeded (sealref_create), so every bootstrap PRNG starts from {seed, 0, ..., 0}
extern "C" int sealref_encrypt_zero_symmetric(sealref_ctx *c, size_t L, uint64_t *out2)
{
    REF_TRY
    Encryptor encryptor(*c->context, c->keygen->secret_key());
    Ciphertext ct;
    encryptor.encrypt_zero_symmetric(level(c, L)->parms_id(), ct);
    store_ct(c, ct, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_bfv_encrypt(sealref_ctx *c, const uint64_t *slots, uint64_t *out2)
{
    REF_TRY
    BatchEncoder enc(*c->context);
    Plaintext pt;
    enc.encode(std::vector<uint64_t>(slots, slots + c->n), pt);
    Encryptor encryptor(*c->context, c->keygen->secret_key());
    Ciphertext ct;
    encryptor.encrypt_symmetric(pt, ct);
    store_ct(c, ct, out2);
    return 0;
    REF_CATCH(-1)
}

extern "C" int sealref_bfv_decrypt(
    sealref_ctx *c, size_t L, size_t size, const uint64_t *ctdata, uint64_t *slots, int *noise_budget)
{
    REF_TRY
    Ciphertext ct = make_ct(c, L, size, ctdata);
    Decryptor dec(*c->context, c->keygen->secret_key());
    if (noise_budget)
        *noise_budget = dec.invariant_noise_budget(ct);
    Plaintext pt;
    dec.decrypt(ct, pt);
    BatchEncoder enc(*c->context);
    std::vector<uint64_t> v;
    enc.decode(pt, v);
    std::memcpy(slots, v.data(), c->n * sizeof(uint64_t));
    return 0;
    REF_CATCH(-1)
}

static void randomize(const sealref_ctx *c, size_t L, Ciphertext &ct, std::mt19937_64 &rng)
{
    auto &mods = level(c, L)->parms().coeff_modulus();
    for (size_t p = 0; p < ct.size(); p++)
        for (size_t i = 0; i < L; i++)
        {
            uint64_t q = mods[i].value();
            uint64_t *row = ct.data(p) + i * c->n;
            for (size_t j = 0; j < c->n; j++)
                row[j] = rng() % q;
        }
}

extern "C" double sealref_time_op(sealref_ctx *c, int op, size_t L, int threads, int reps)
{
    REF_TRY
    const RelinKeys *rk = nullptr;
    const GaloisKeys *gk = nullptr;
    uint32_t elt = 0;
    if (op == 0)
        rk = &relin_keys(c);
    if (op == 2)
    {
        elt = c->context->key_context_data()->galois_tool()->get_elt_from_step(1);
        gk = &galois_for(c, elt);
    }
    std::vector<std::string> errs(static_cast<size_t>(threads));
    // inputs prepared outside the timed region (as native/bench does with PauseTiming)
    std::vector<std::vector<Ciphertext>> as(threads), bs(threads);
    std::vector<MemoryPoolHandle> pools;
    for (int t = 0; t < threads; t++)
    {
        pools.push_back(MemoryPoolHandle::New());
        std::mt19937_64 rng(0x5EA1 + t);
        as[t].push_back(make_ct(c, L, 2, nullptr, pools[t]));
        bs[t].push_back(make_ct(c, L, 2, nullptr, pools[t]));
        randomize(c, L, as[t][0], rng);
        randomize(c, L, bs[t][0], rng);
        if (op == 3)
            as[t][0].scale() = static_cast<double>(level(c, L)->parms().coeff_modulus().back().value()) * 1024.0;
    }
    auto worker = [&](int t) {
        try
        {
            for (int r = 0; r < reps; r++)
            {
                Ciphertext x(pools[t]);
                x = as[t][0];
                switch (op)
                {
                case 0:
                    c->evaluator->multiply_inplace(x, bs[t][0], pools[t]);
                    c->evaluator->relinearize_inplace(x, *rk, pools[t]);
                    break;
                case 1:
                    if (c->scheme == scheme_type::bfv)
                    {
                        c->evaluator->transform_to_ntt_inplace(x);
                        c->evaluator->transform_from_ntt_inplace(x);
                    }
                    else
                    {
                        c->evaluator->transform_from_ntt_inplace(x);
                        c->evaluator->transform_to_ntt_inplace(x);
                    }
                    break;
                case 2:
                    c->evaluator->apply_galois_inplace(x, elt, *gk, pools[t]);
                    break;
                case 3:
                    c->evaluator->rescale_to_next_inplace(x, pools[t]);
                    break;
                case 4:
                    c->evaluator->multiply_inplace(x, bs[t][0], pools[t]);
                    break;
                default:
                    throw std::invalid_argument("unknown op");
                }
            }
        }
        catch (const std::exception &e)
        {
            errs[t] = e.what();
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back(worker, t);
    for (auto &x : th)
        x.join();
    auto t1 = std::chrono::steady_clock::now();
    for (auto &e : errs)
        if (!e.empty())
        {
            g_err = e;
            return -1.0;
        }
    return std::chrono::duration<double>(t1 - t0).count();
    REF_CATCH(-1.0)
}
