// include/seal_b200/evaluator.hpp -- header-only C++17 drop-in for the hot-path members of seal::Evaluator.
//
// seal::Evaluator has no virtual members and cannot be subclassed (evaluator.h:79, copy/move deleted :1320-1326), so the
// drop-in is a class with the SAME member signatures (same seal::Ciphertext / RelinKeys / GaloisKeys / SEALContext
// types, same argument meaning, same exception types from the same prologue checks) that a SEAL user instantiates
// instead of seal::Evaluator for multiply / relinearize / rescale_to_next / mod_switch_to_next / apply_galois /
// rotate_rows / rotate_columns / rotate_vector / complex_conjugate / transform_{to,from}_ntt.  It is compiled against the
// user's own SEAL headers; all arithmetic happens in libseal_b200.so (include/seal_b200.h) on the GPU -- there is
// no CPU fallback.  Batch overloads (std::vector<Ciphertext>) are the extension the reference does not have.
//
// Reference members mirrored (native/src/seal/): evaluator.h:219-345, 510-545, 929-1316; evaluator.cpp:352-393, 569-708,
// 1144-1199, 1201-1294, 1404-1541, 2289-2559.
#pragma once
#include "../seal_b200.h"
#include "seal/seal.h"
#include "seal/util/numth.h"
#include "seal/util/uintarithsmallmod.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace seal_b200
{
    class Evaluator;

    // ---- device-resident ciphertexts ---------------------------------------------------------------------------------
    // A batch of B ciphertexts of one shape that stays on the GPU between Evaluator calls.  It carries exactly the metadata
    // seal::Ciphertext carries (ciphertext.h:728-742: parms_id, is_ntt_form, size, poly_modulus_degree, coeff_modulus_size,
    // scale, correction_factor), shared by all members of the batch; Evaluator updates it the way the reference updates the
    // members of each Ciphertext.  Storage is one slab [B][size][L][n] laid out like Ciphertext::data(), allocated from the
    // Evaluator's context (sb200_device_malloc) -- the reference's counterpart is the MemoryPoolHandle behind a Ciphertext.
    // Usage: ev.upload(cts, batch); ev.multiply_relinearize_inplace(batch, other, rk); ev.rescale_to_next_inplace(batch); ...
    //        ev.download(batch, cts);      -- one H2D and one D2H for a whole chain of operations.
    class CiphertextBatch
    {
    public:
        CiphertextBatch() = default;
        ~CiphertextBatch() { release(); }
        CiphertextBatch(const CiphertextBatch &) = delete;
        CiphertextBatch &operator=(const CiphertextBatch &) = delete;
        CiphertextBatch(CiphertextBatch &&o) noexcept { *this = std::move(o); }
        CiphertextBatch &operator=(CiphertextBatch &&o) noexcept
        {
            if (this != &o)
            {
                release();
                ctx_ = o.ctx_, d_ = o.d_, cap_ = o.cap_, batch_ = o.batch_, size_ = o.size_, L_ = o.L_, n_ = o.n_;
                parms_id_ = o.parms_id_, ntt_ = o.ntt_, scale_ = o.scale_, cf_ = o.cf_;
                o.ctx_ = nullptr, o.d_ = nullptr, o.cap_ = o.batch_ = o.size_ = o.L_ = o.n_ = 0;
            }
            return *this;
        }
        std::size_t batch_size() const noexcept { return batch_; }
        std::size_t size() const noexcept { return size_; }                             // polynomials per ciphertext
        std::size_t coeff_modulus_size() const noexcept { return L_; }
        std::size_t poly_modulus_degree() const noexcept { return n_; }
        const seal::parms_id_type &parms_id() const noexcept { return parms_id_; }
        bool is_ntt_form() const noexcept { return ntt_; }
        double &scale() noexcept { return scale_; }
        double scale() const noexcept { return scale_; }
        std::uint64_t correction_factor() const noexcept { return cf_; }
        std::uint64_t *device_data() noexcept { return d_; }
        const std::uint64_t *device_data() const noexcept { return d_; }
        std::size_t words_per_ciphertext() const noexcept { return size_ * L_ * n_; }
        void release() noexcept
        {
            if (d_ && ctx_)
                sb200_device_free(ctx_, d_);
            d_ = nullptr, cap_ = 0;
        }

    private:
        friend class Evaluator;
        sb200_context *ctx_ = nullptr;
        std::uint64_t *d_ = nullptr;
        std::size_t cap_ = 0; // words
        std::size_t batch_ = 0, size_ = 0, L_ = 0, n_ = 0;
        seal::parms_id_type parms_id_ = seal::parms_id_zero;
        bool ntt_ = false;
        double scale_ = 1.0;
        std::uint64_t cf_ = 1;
    };

    class Evaluator
    {
    public:
        explicit Evaluator(const seal::SEALContext &context, int device = 0) : context_(context)
        {
            if (!context_.parameters_set())
                throw std::invalid_argument("encryption parameters are not set correctly"); // evaluator.cpp:121-128
            auto &parms = context_.key_context_data()->parms();
            std::vector<std::uint64_t> q;
            for (auto &m : parms.coeff_modulus())
                q.push_back(m.value());
            scheme_ = parms.scheme();
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::ckks && scheme_ != seal::scheme_type::bgv)
                throw std::invalid_argument("unsupported scheme");
            check(sb200_context_create(
                static_cast<int>(scheme_), parms.poly_modulus_degree(), q.data(), q.size(),
                scheme_ != seal::scheme_type::ckks ? parms.plain_modulus().value() : 0, device, &ctx_));
        }
        ~Evaluator()
        {
            for (auto &kv : keys_)
                sb200_kswitch_key_destroy(kv.second.handle);
            tmp_.release(); // before the context it was allocated from goes away
            if (ctx_)
                sb200_context_destroy(ctx_);
        }
        Evaluator(const Evaluator &) = delete;
        Evaluator &operator=(const Evaluator &) = delete;

        // ---- multiply (evaluator.cpp:352-393) ------------------------------------------------------------------
        void multiply_inplace(seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2,
                              seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted1, "encrypted1 is not valid for encryption parameters");
            validate(encrypted2, "encrypted2 is not valid for encryption parameters");
            if (encrypted1.parms_id() != encrypted2.parms_id())
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            const bool ckks = scheme_ == seal::scheme_type::ckks, bgv = scheme_ == seal::scheme_type::bgv;
            if ((ckks || bgv) && !(encrypted1.is_ntt_form() && encrypted2.is_ntt_form()))
                throw std::invalid_argument("encrypted1 or encrypted2 must be in NTT form"); // :571-574, :712-715
            if (!(ckks || bgv) && (encrypted1.is_ntt_form() || encrypted2.is_ntt_form()))
                throw std::invalid_argument("encrypted1 or encrypted2 cannot be in NTT form"); // :397-400
            auto cd = context_.get_context_data(encrypted1.parms_id());
            const std::size_t L = encrypted1.coeff_modulus_size(), n = encrypted1.poly_modulus_degree();
            const std::size_t s1 = encrypted1.size(), s2 = encrypted2.size(); // any sizes: :524-560, :664-700, :796-833
            std::vector<std::uint64_t> a(encrypted1.data(), encrypted1.data() + s1 * L * n);
            double new_scale = encrypted1.scale() * encrypted2.scale();
            if (ckks && !scale_within_bounds(new_scale, *cd))
                throw std::invalid_argument("scale out of bounds"); // :703-707
            encrypted1.resize(context_, cd->parms_id(), s1 + s2 - 1); // throws on > SEAL_CIPHERTEXT_SIZE_MAX like the reference
            check(sb200_multiply_sized_host(ctx_, L, s1, s2, 1, a.data(), encrypted2.data(), encrypted1.data()));
            if (ckks)
                encrypted1.scale() = new_scale;
            if (bgv) // :838-840
                encrypted1.correction_factor() = seal::util::multiply_uint_mod(
                    encrypted1.correction_factor(), encrypted2.correction_factor(), cd->parms().plain_modulus());
            throw_if_transparent(encrypted1);
        }
        void multiply(const seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, seal::Ciphertext &destination,
                      seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (&encrypted2 == &destination)
            {
                multiply_inplace(destination, encrypted1, std::move(pool));
            }
            else
            {
                destination = encrypted1;
                multiply_inplace(destination, encrypted2, std::move(pool));
            }
        }

        // ---- square (evaluator.cpp:843-1142): same residues as multiply(x, x) -------------------------------------------
        void square_inplace(seal::Ciphertext &encrypted, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            seal::Ciphertext copy = encrypted;
            multiply_inplace(encrypted, copy, std::move(pool));
        }
        void square(const seal::Ciphertext &encrypted, seal::Ciphertext &destination,
                    seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            square_inplace(destination, std::move(pool));
        }

        // ---- negate / add / sub (evaluator.cpp:130-350) ------------------------------------------------------------------
        void negate_inplace(seal::Ciphertext &encrypted) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            check(sb200_negate_host(ctx_, encrypted.coeff_modulus_size(), encrypted.size(), 1, encrypted.data(), encrypted.data()));
            throw_if_transparent(encrypted);
        }
        void negate(const seal::Ciphertext &encrypted, seal::Ciphertext &destination) const
        {
            destination = encrypted;
            negate_inplace(destination);
        }
        void add_inplace(seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2) const { linear(encrypted1, encrypted2, false); }
        void add(const seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, seal::Ciphertext &destination) const
        {
            if (&encrypted2 == &destination)
            {
                add_inplace(destination, encrypted1);
            }
            else
            {
                destination = encrypted1;
                add_inplace(destination, encrypted2);
            }
        }
        void sub_inplace(seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2) const { linear(encrypted1, encrypted2, true); }
        void sub(const seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, seal::Ciphertext &destination) const
        {
            if (&encrypted2 == &destination)
            {
                sub_inplace(destination, encrypted1);
                negate_inplace(destination);
            }
            else
            {
                destination = encrypted1;
                sub_inplace(destination, encrypted2);
            }
        }

        // ---- multiply_plain (evaluator.cpp:1975-2019 dispatcher, :2157-2195 multiply_plain_ntt, :2021-2155 multiply_plain_normal) ----
        void multiply_plain_inplace(seal::Ciphertext &encrypted, const seal::Plaintext &plain,
                                    seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (!seal::is_metadata_valid_for(plain, context_) || !seal::is_buffer_valid(plain))
                throw std::invalid_argument("plain is not valid for encryption parameters");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            if (!plain.is_ntt_form())
            {
                // multiply_plain_normal (:2021-2155) / transform + multiply_plain_ntt (:1999-2004)
                if (scheme_ == seal::scheme_type::ckks)
                    throw std::invalid_argument("plain is not valid for encryption parameters"); // CKKS plaintexts are always in NTT form
                const std::vector<std::uint64_t> words = padded(plain);
                check(sb200_multiply_plain_coeff_host(ctx_, encrypted.coeff_modulus_size(), encrypted.size(), 1, encrypted.is_ntt_form() ? 1 : 0,
                                                      encrypted.data(), words.data(), encrypted.data()));
                throw_if_transparent(encrypted);
                return;
            }
            const bool back = !encrypted.is_ntt_form(); // :2006-2011: to NTT form, multiply, back
            if (back)
                transform_to_ntt_inplace(encrypted);
            if (encrypted.parms_id() != plain.parms_id())
                throw std::invalid_argument("encrypted_ntt and plain_ntt parameter mismatch"); // :2164-2167
            check(sb200_multiply_plain_host(ctx_, encrypted.coeff_modulus_size(), encrypted.size(), 1, encrypted.data(), plain.data(),
                                            encrypted.data()));
            encrypted.scale() *= plain.scale();
            if (!scale_within_bounds(encrypted.scale(), *context_.get_context_data(encrypted.parms_id())))
                throw std::invalid_argument("scale out of bounds"); // :2189-2193
            if (back)
                transform_from_ntt_inplace(encrypted);
            throw_if_transparent(encrypted);
        }
        void multiply_plain(const seal::Ciphertext &encrypted, const seal::Plaintext &plain, seal::Ciphertext &destination,
                            seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            multiply_plain_inplace(destination, plain, std::move(pool));
        }

        // ---- add_plain / sub_plain (evaluator.cpp:1759-1868, :1870-1975) ---------------------------------------------------
        void add_plain_inplace(seal::Ciphertext &encrypted, const seal::Plaintext &plain,
                               seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            plain_linear(encrypted, plain, false);
            (void)pool;
        }
        void add_plain(const seal::Ciphertext &encrypted, const seal::Plaintext &plain, seal::Ciphertext &destination,
                       seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            add_plain_inplace(destination, plain, std::move(pool));
        }
        void sub_plain_inplace(seal::Ciphertext &encrypted, const seal::Plaintext &plain,
                               seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            plain_linear(encrypted, plain, true);
            (void)pool;
        }
        void sub_plain(const seal::Ciphertext &encrypted, const seal::Plaintext &plain, seal::Ciphertext &destination,
                       seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            sub_plain_inplace(destination, plain, std::move(pool));
        }

        // ---- compositions the reference builds from the operations above ---------------------------------------------
        // add_many (evaluator.cpp:236-262)
        void add_many(const std::vector<seal::Ciphertext> &encrypteds, seal::Ciphertext &destination) const
        {
            if (encrypteds.empty())
                throw std::invalid_argument("encrypteds cannot be empty");
            for (auto &e : encrypteds)
                if (&e == &destination)
                    throw std::invalid_argument("encrypteds must be different from destination");
            destination = encrypteds[0];
            for (std::size_t i = 1; i < encrypteds.size(); i++)
                add_inplace(destination, encrypteds[i]);
        }
        // multiply_many (evaluator.cpp:1649-1724): pairwise products, each relinearized, appended until one is left
        void multiply_many(const std::vector<seal::Ciphertext> &encrypteds, const seal::RelinKeys &relin_keys, seal::Ciphertext &destination,
                           seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (encrypteds.empty())
                throw std::invalid_argument("encrypteds vector must not be empty");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            for (auto &e : encrypteds)
                if (&e == &destination)
                    throw std::invalid_argument("encrypteds must be different from destination");
            if (!context_.get_context_data(encrypteds[0].parms_id()))
                throw std::invalid_argument("encrypteds is not valid for encryption parameters");
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::bgv)
                throw std::logic_error("unsupported scheme");
            if (encrypteds.size() == 1)
            {
                destination = encrypteds[0];
                return;
            }
            std::vector<seal::Ciphertext> products;
            for (std::size_t i = 0; i + 1 < encrypteds.size(); i += 2)
            {
                seal::Ciphertext temp;
                if (encrypteds[i].data() == encrypteds[i + 1].data())
                    square(encrypteds[i], temp);
                else
                    multiply(encrypteds[i], encrypteds[i + 1], temp);
                relinearize_inplace(temp, relin_keys, pool);
                products.emplace_back(std::move(temp));
            }
            if (encrypteds.size() & 1)
                products.emplace_back(encrypteds.back());
            for (std::size_t i = 0; i + 1 < products.size(); i += 2)
            {
                seal::Ciphertext temp;
                multiply(products[i], products[i + 1], temp);
                relinearize_inplace(temp, relin_keys, pool);
                products.emplace_back(std::move(temp));
            }
            destination = products.back();
        }
        // exponentiate (evaluator.cpp:1726-1757)
        void exponentiate_inplace(seal::Ciphertext &encrypted, std::uint64_t exponent, const seal::RelinKeys &relin_keys,
                                  seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (!context_.get_context_data(encrypted.parms_id()))
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (!context_.get_context_data(relin_keys.parms_id()))
                throw std::invalid_argument("relin_keys is not valid for encryption parameters");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            if (exponent == 0)
                throw std::invalid_argument("exponent cannot be 0");
            if (exponent == 1)
                return;
            std::vector<seal::Ciphertext> copies(static_cast<std::size_t>(exponent), encrypted);
            multiply_many(copies, relin_keys, encrypted, std::move(pool));
        }
        void exponentiate(const seal::Ciphertext &encrypted, std::uint64_t exponent, const seal::RelinKeys &relin_keys,
                          seal::Ciphertext &destination, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            exponentiate_inplace(destination, exponent, relin_keys, std::move(pool));
        }
        // mod_switch_to / rescale_to (evaluator.cpp:1451-1473, :1543-1590): repeat the single step down to parms_id
        void mod_switch_to_inplace(seal::Ciphertext &encrypted, seal::parms_id_type parms_id,
                                   seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            check_target_level(encrypted, parms_id);
            while (encrypted.parms_id() != parms_id)
                mod_switch_to_next_inplace(encrypted, pool);
        }
        void mod_switch_to(const seal::Ciphertext &encrypted, seal::parms_id_type parms_id, seal::Ciphertext &destination,
                           seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            mod_switch_to_inplace(destination, parms_id, std::move(pool));
        }
        void rescale_to_inplace(seal::Ciphertext &encrypted, seal::parms_id_type parms_id,
                                seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            check_target_level(encrypted, parms_id);
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            if (scheme_ != seal::scheme_type::ckks)
                throw std::invalid_argument("unsupported operation for scheme type");
            while (encrypted.parms_id() != parms_id)
            {
                seal::Ciphertext next;
                mod_switch_impl(encrypted, next, true);
                encrypted = std::move(next);
            }
        }
        void rescale_to(const seal::Ciphertext &encrypted, seal::parms_id_type parms_id, seal::Ciphertext &destination,
                        seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            rescale_to_inplace(destination, parms_id, std::move(pool));
        }

        // ---- dropping RNS components without scaling: layout only, no arithmetic ----------------------------------------
        // mod_reduce_to_next / mod_reduce_to (evaluator.cpp:1598-1647 -> mod_switch_drop_to_next :1296-1366)
        void mod_reduce_to_next_inplace(seal::Ciphertext &encrypted, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (context_.last_parms_id() == encrypted.parms_id())
                throw std::invalid_argument("end of modulus switching chain reached");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            drop_last_component(encrypted);
            throw_if_transparent(encrypted);
        }
        void mod_reduce_to_next(const seal::Ciphertext &encrypted, seal::Ciphertext &destination,
                                seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            mod_reduce_to_next_inplace(destination, std::move(pool));
        }
        void mod_reduce_to_inplace(seal::Ciphertext &encrypted, seal::parms_id_type parms_id,
                                   seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            check_target_level(encrypted, parms_id);
            while (encrypted.parms_id() != parms_id)
                mod_reduce_to_next_inplace(encrypted, pool);
        }
        void mod_reduce_to(const seal::Ciphertext &encrypted, seal::parms_id_type parms_id, seal::Ciphertext &destination,
                           seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            mod_reduce_to_inplace(destination, parms_id, std::move(pool));
        }
        // NTT-form plaintexts one level down (evaluator.h:431-458, evaluator.cpp:1368-1402): the last component is cut off
        void mod_switch_to_next_inplace(seal::Plaintext &plain) const
        {
            if (!seal::is_valid_for(plain, context_))
                throw std::invalid_argument("plain is not valid for encryption parameters");
            auto cd = context_.get_context_data(plain.parms_id());
            if (!plain.is_ntt_form())
                throw std::invalid_argument("plain is not in NTT form");
            if (!cd->next_context_data())
                throw std::invalid_argument("end of modulus switching chain reached");
            auto &next = *cd->next_context_data();
            if (!scale_within_bounds(plain.scale(), next))
                throw std::invalid_argument("scale out of bounds");
            const std::size_t words = next.parms().coeff_modulus().size() * next.parms().poly_modulus_degree();
            plain.parms_id() = seal::parms_id_zero;
            plain.resize(words);
            plain.parms_id() = next.parms_id();
        }
        void mod_switch_to_next(const seal::Plaintext &plain, seal::Plaintext &destination) const
        {
            destination = plain;
            mod_switch_to_next_inplace(destination);
        }
        void mod_switch_to_inplace(seal::Plaintext &plain, seal::parms_id_type parms_id) const
        {
            auto cur = context_.get_context_data(plain.parms_id());
            auto target = context_.get_context_data(parms_id);
            if (!cur)
                throw std::invalid_argument("plain is not valid for encryption parameters");
            if (!target)
                throw std::invalid_argument("parms_id is not valid for encryption parameters");
            if (!plain.is_ntt_form())
                throw std::invalid_argument("plain is not in NTT form");
            if (cur->chain_index() < target->chain_index())
                throw std::invalid_argument("cannot switch to higher level modulus");
            while (plain.parms_id() != parms_id)
                mod_switch_to_next_inplace(plain);
        }
        void mod_switch_to(const seal::Plaintext &plain, seal::parms_id_type parms_id, seal::Plaintext &destination) const
        {
            destination = plain;
            mod_switch_to_inplace(destination, parms_id);
        }

        // ---- relinearize (evaluator.cpp:1144-1199) ---------------------------------------------------------------
        void relinearize_inplace(seal::Ciphertext &encrypted, const seal::RelinKeys &relin_keys,
                                 seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            auto cd = context_.get_context_data(encrypted.parms_id());
            if (!cd)
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (relin_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("relin_keys is not valid for encryption parameters");
            if (encrypted.size() < 2)
                throw std::invalid_argument("destination_size must be at least 2 and less than or equal to current count");
            if (relin_keys.size() < encrypted.size() - 2)
                throw std::invalid_argument("not enough relinearization keys");
            if (encrypted.size() == 2)
                return;
            check_keyswitch_operand(encrypted, pool);
            const std::size_t L = encrypted.coeff_modulus_size(), n = encrypted.poly_modulus_degree(), size = encrypted.size();
            std::lock_guard<std::mutex> lock(mu_);
            if (size == 3)
            {
                sb200_kswitch_key *key = key_for(relin_keys, seal::RelinKeys::get_index(2), L);
                std::vector<std::uint64_t> in(encrypted.data(), encrypted.data() + 3 * L * n);
                encrypted.resize(context_, cd->parms_id(), 2);
                check(sb200_relinearize_host(ctx_, L, 1, in.data(), key, encrypted.data()));
                throw_if_transparent(encrypted);
                return;
            }
            // size > 3, exactly as the reference's loop is written (evaluator.cpp:1176-1187): every step key-switches the LAST
            // polynomial (the iterator is not advanced) with the key of index size-1-I, then the ciphertext is cut to 2 polynomials
            std::vector<std::uint64_t> cur(encrypted.data(), encrypted.data() + size * L * n), next(cur.size());
            for (std::size_t I = 0; I < size - 2; I++)
            {
                sb200_kswitch_key *key = key_for(relin_keys, seal::RelinKeys::get_index(size - 1 - I), L);
                check(sb200_relinearize_sized_host(ctx_, L, size, 1, cur.data(), key, next.data()));
                cur.swap(next);
            }
            encrypted.resize(context_, cd->parms_id(), 2);
            std::memcpy(encrypted.data(), cur.data(), 2 * L * n * sizeof(std::uint64_t));
            throw_if_transparent(encrypted);
        }
        void relinearize(const seal::Ciphertext &encrypted, const seal::RelinKeys &relin_keys, seal::Ciphertext &destination,
                         seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            relinearize_inplace(destination, relin_keys, std::move(pool));
        }

        // ---- rescale / mod switch (evaluator.cpp:1404-1541, 1201-1358) --------------------------------------------
        void rescale_to_next(const seal::Ciphertext &encrypted, seal::Ciphertext &destination,
                             seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (context_.last_parms_id() == encrypted.parms_id())
                throw std::invalid_argument("end of modulus switching chain reached");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            if (scheme_ != seal::scheme_type::ckks)
                throw std::invalid_argument("unsupported operation for scheme type");
            mod_switch_impl(encrypted, destination, true);
        }
        void rescale_to_next_inplace(seal::Ciphertext &encrypted, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            rescale_to_next(encrypted, encrypted, std::move(pool));
        }
        void mod_switch_to_next(const seal::Ciphertext &encrypted, seal::Ciphertext &destination,
                                seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (context_.last_parms_id() == encrypted.parms_id())
                throw std::invalid_argument("end of modulus switching chain reached");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            mod_switch_impl(encrypted, destination, false);
        }
        void mod_switch_to_next_inplace(seal::Ciphertext &encrypted, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            mod_switch_to_next(encrypted, encrypted, std::move(pool));
        }

        // ---- Galois automorphisms and rotations (evaluator.cpp:2384-2559, evaluator.h:1006-1316) ---------------------
        void apply_galois_inplace(seal::Ciphertext &encrypted, std::uint32_t galois_elt, const seal::GaloisKeys &galois_keys,
                                  seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (galois_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("galois_keys is not valid for encryption parameters");
            if (!galois_keys.has_key(galois_elt))
                throw std::invalid_argument("Galois key not present");
            const std::size_t L = encrypted.coeff_modulus_size(), n = encrypted.poly_modulus_degree();
            if (!(galois_elt & 1) || galois_elt >= 2 * n)
                throw std::invalid_argument("Galois element is not valid");
            if (encrypted.size() != 2)
                throw std::invalid_argument("encrypted size must be 2");
            check_keyswitch_operand(encrypted, pool);
            std::lock_guard<std::mutex> lock(mu_); // the cached key stays valid (not evicted) while the operation uses it
            sb200_kswitch_key *key = key_for(galois_keys, seal::GaloisKeys::get_index(galois_elt), L);
            std::vector<std::uint64_t> in(encrypted.data(), encrypted.data() + 2 * L * n);
            check(sb200_apply_galois_host(ctx_, L, 1, in.data(), galois_elt, key, encrypted.data()));
            throw_if_transparent(encrypted);
        }
        void apply_galois(const seal::Ciphertext &encrypted, std::uint32_t galois_elt, const seal::GaloisKeys &galois_keys,
                          seal::Ciphertext &destination, seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            apply_galois_inplace(destination, galois_elt, galois_keys, std::move(pool));
        }
        void rotate_rows_inplace(seal::Ciphertext &encrypted, int steps, const seal::GaloisKeys &galois_keys,
                                 seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::bgv) // evaluator.h:1076-1080, 1145-1149
                throw std::logic_error("unsupported scheme");
            rotate_internal(encrypted, steps, galois_keys, std::move(pool));
        }
        void rotate_rows(const seal::Ciphertext &encrypted, int steps, const seal::GaloisKeys &galois_keys, seal::Ciphertext &destination,
                         seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            rotate_rows_inplace(destination, steps, galois_keys, std::move(pool));
        }
        void rotate_columns_inplace(seal::Ciphertext &encrypted, const seal::GaloisKeys &galois_keys,
                                    seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::bgv) // evaluator.h:1076-1080, 1145-1149
                throw std::logic_error("unsupported scheme");
            conjugate_internal(encrypted, galois_keys, std::move(pool));
        }
        void rotate_columns(const seal::Ciphertext &encrypted, const seal::GaloisKeys &galois_keys, seal::Ciphertext &destination,
                            seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            rotate_columns_inplace(destination, galois_keys, std::move(pool));
        }
        void rotate_vector_inplace(seal::Ciphertext &encrypted, int steps, const seal::GaloisKeys &galois_keys,
                                   seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (scheme_ != seal::scheme_type::ckks)
                throw std::logic_error("unsupported scheme");
            rotate_internal(encrypted, steps, galois_keys, std::move(pool));
        }
        void rotate_vector(const seal::Ciphertext &encrypted, int steps, const seal::GaloisKeys &galois_keys, seal::Ciphertext &destination,
                           seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            rotate_vector_inplace(destination, steps, galois_keys, std::move(pool));
        }
        void complex_conjugate_inplace(seal::Ciphertext &encrypted, const seal::GaloisKeys &galois_keys,
                                       seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (scheme_ != seal::scheme_type::ckks)
                throw std::logic_error("unsupported scheme");
            conjugate_internal(encrypted, galois_keys, std::move(pool));
        }
        void complex_conjugate(const seal::Ciphertext &encrypted, const seal::GaloisKeys &galois_keys, seal::Ciphertext &destination,
                               seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination = encrypted;
            complex_conjugate_inplace(destination, galois_keys, std::move(pool));
        }

        // ---- NTT form changes (evaluator.cpp:2289-2382) -------------------------------------------------------------
        void transform_to_ntt_inplace(seal::Ciphertext &encrypted) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (encrypted.is_ntt_form())
                throw std::invalid_argument("encrypted is already in NTT form");
            check(sb200_ntt_forward_host(ctx_, encrypted.coeff_modulus_size(), encrypted.size(), 1, encrypted.data()));
            encrypted.is_ntt_form() = true;
            throw_if_transparent(encrypted);
        }
        void transform_to_ntt(const seal::Ciphertext &encrypted, seal::Ciphertext &destination_ntt) const
        {
            destination_ntt = encrypted;
            transform_to_ntt_inplace(destination_ntt);
        }
        void transform_from_ntt_inplace(seal::Ciphertext &encrypted_ntt) const
        {
            validate(encrypted_ntt, "encrypted is not valid for encryption parameters");
            if (!encrypted_ntt.is_ntt_form())
                throw std::invalid_argument("encrypted_ntt is not in NTT form");
            check(sb200_ntt_inverse_host(ctx_, encrypted_ntt.coeff_modulus_size(), encrypted_ntt.size(), 1, encrypted_ntt.data()));
            encrypted_ntt.is_ntt_form() = false;
            throw_if_transparent(encrypted_ntt);
        }
        void transform_from_ntt(const seal::Ciphertext &encrypted_ntt, seal::Ciphertext &destination) const
        {
            destination = encrypted_ntt;
            transform_from_ntt_inplace(destination);
        }

        // transform_to_ntt_inplace(Plaintext&, parms_id) (evaluator.cpp:2197-2287)
        void transform_to_ntt_inplace(seal::Plaintext &plain, seal::parms_id_type parms_id,
                                      seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            if (!seal::is_valid_for(plain, context_))
                throw std::invalid_argument("plain is not valid for encryption parameters");
            auto cd = context_.get_context_data(parms_id);
            if (!cd)
                throw std::invalid_argument("parms_id is not valid for the current context");
            if (plain.is_ntt_form())
                throw std::invalid_argument("plain is already in NTT form");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            const std::vector<std::uint64_t> words = padded(plain);
            const std::size_t L = cd->parms().coeff_modulus().size(), n = cd->parms().poly_modulus_degree();
            plain.resize(n * L);
            check(sb200_plain_to_ntt_host(ctx_, L, 1, words.data(), plain.data()));
            plain.parms_id() = parms_id;
        }
        void transform_to_ntt(const seal::Plaintext &plain, seal::parms_id_type parms_id, seal::Plaintext &destination_ntt,
                              seal::MemoryPoolHandle pool = seal::MemoryManager::GetPool()) const
        {
            destination_ntt = plain;
            transform_to_ntt_inplace(destination_ntt, parms_id, std::move(pool));
        }

        // ---- batch extension: destination[i] = relinearize(multiply(a[i], b[i])), one device pass over the batch ----
        // (upload through page-locked staging, fused multiply + relinearize on the device, download).  Scales and BGV correction
        // factors may differ from ciphertext to ciphertext; they are computed before anything is written, so destination may
        // alias a or b.
        void multiply_relinearize(const std::vector<seal::Ciphertext> &a, const std::vector<seal::Ciphertext> &b,
                                  const seal::RelinKeys &relin_keys, std::vector<seal::Ciphertext> &destination) const
        {
            if (a.size() != b.size() || a.empty())
                throw std::invalid_argument("batch size mismatch");
            if (relin_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("relin_keys is not valid for encryption parameters");
            const std::size_t B = a.size();
            const bool ckks = scheme_ == seal::scheme_type::ckks, ntt = scheme_ != seal::scheme_type::bfv;
            auto cd = context_.get_context_data(a[0].parms_id());
            std::vector<double> scales(B);
            std::vector<std::uint64_t> factors(B);
            for (std::size_t i = 0; i < B; i++)
            {
                validate(a[i], "encrypted1 is not valid for encryption parameters");
                validate(b[i], "encrypted2 is not valid for encryption parameters");
                if (a[i].parms_id() != a[0].parms_id() || b[i].parms_id() != a[0].parms_id())
                    throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
                if (a[i].size() != 2 || b[i].size() != 2 || a[i].is_ntt_form() != ntt || b[i].is_ntt_form() != ntt)
                    throw std::invalid_argument("batch entries must be fresh size-2 ciphertexts in the scheme's native form");
                scales[i] = ckks ? a[i].scale() * b[i].scale() : a[i].scale(); // evaluator.cpp:703-707; BFV / BGV leave the scale alone
                if (ckks && !scale_within_bounds(scales[i], *cd))
                    throw std::invalid_argument("scale out of bounds");
                factors[i] = scheme_ == seal::scheme_type::bgv
                                 ? seal::util::multiply_uint_mod(a[i].correction_factor(), b[i].correction_factor(), cd->parms().plain_modulus())
                                 : a[i].correction_factor();
            }
            CiphertextBatch da, db;
            upload_rows(a.data(), B, da);
            upload_rows(b.data(), B, db);
            da.scale_ = db.scale_ = 1.0; // per-ciphertext metadata is applied below
            multiply_relinearize_inplace(da, db, relin_keys);
            download(da, destination);
            for (std::size_t i = 0; i < B; i++)
            {
                destination[i].scale() = scales[i];
                destination[i].correction_factor() = factors[i];
            }
        }

        // =================================== device-resident batches (CiphertextBatch) ===================================
        // Same members, same checks and metadata updates as the seal::Ciphertext overloads above; the data never leaves the GPU.
        void upload(const std::vector<seal::Ciphertext> &cts, CiphertextBatch &destination) const
        {
            upload(cts.data(), cts.size(), destination);
        }
        // a contiguous range of ciphertext objects (e.g. one slice of a larger vector: slices of a batch can be uploaded, processed and
        // downloaded by different host threads so that the copies of one slice overlap the kernels of another)
        void upload(const seal::Ciphertext *first, std::size_t count, CiphertextBatch &destination) const
        {
            if (!first || !count)
                throw std::invalid_argument("batch cannot be empty");
            const seal::Ciphertext &f = first[0];
            for (std::size_t i = 0; i < count; i++)
            {
                const seal::Ciphertext &c = first[i];
                validate(c, "encrypted is not valid for encryption parameters");
                const double s1 = c.scale(), s2 = f.scale();
                if (c.parms_id() != f.parms_id() || c.size() != f.size() || c.is_ntt_form() != f.is_ntt_form() ||
                    c.correction_factor() != f.correction_factor() || std::memcmp(&s1, &s2, sizeof(double)) != 0)
                    throw std::invalid_argument("batch members must share parms_id, size, NTT form, scale and correction factor");
            }
            upload_rows(first, count, destination);
        }
        void download(const CiphertextBatch &source, std::vector<seal::Ciphertext> &cts) const
        {
            cts.resize(source.batch_);
            download(source, cts.data());
        }
        void download(const CiphertextBatch &source, seal::Ciphertext *cts) const
        {
            owned(source);
            if (!cts)
                throw std::invalid_argument("destination cannot be null");
            std::vector<std::uint64_t *> rows(source.batch_);
            for (std::size_t i = 0; i < source.batch_; i++)
            {
                cts[i].resize(context_, source.parms_id_, source.size_);
                cts[i].is_ntt_form() = source.ntt_;
                cts[i].scale() = source.scale_;
                cts[i].correction_factor() = source.cf_;
                rows[i] = cts[i].data();
            }
            check(sb200_download_rows(ctx_, rows.data(), source.d_, source.words_per_ciphertext() * sizeof(std::uint64_t), source.batch_));
            for (std::size_t i = 0; i < source.batch_; i++)
                throw_if_transparent(cts[i]);
        }
        void copy(const CiphertextBatch &source, CiphertextBatch &destination) const
        {
            owned(source);
            std::lock_guard<std::mutex> lock(mu_);
            shape(destination, source.batch_, source.size_, source.L_, source.n_);
            copy_meta(source, destination);
            check(sb200_memcpy_d2d(ctx_, destination.d_, source.d_, source.batch_ * source.words_per_ciphertext() * sizeof(std::uint64_t), nullptr));
        }
        // Evaluator::multiply_inplace (evaluator.cpp:352-393), any sizes
        void multiply_inplace(CiphertextBatch &encrypted1, const CiphertextBatch &encrypted2) const
        {
            owned(encrypted1), owned(encrypted2);
            if (encrypted1.parms_id_ != encrypted2.parms_id_ || encrypted1.batch_ != encrypted2.batch_)
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            const bool ckks = scheme_ == seal::scheme_type::ckks, bgv = scheme_ == seal::scheme_type::bgv;
            if ((ckks || bgv) && !(encrypted1.ntt_ && encrypted2.ntt_))
                throw std::invalid_argument("encrypted1 or encrypted2 must be in NTT form");
            if (!(ckks || bgv) && (encrypted1.ntt_ || encrypted2.ntt_))
                throw std::invalid_argument("encrypted1 or encrypted2 cannot be in NTT form");
            auto cd = context_.get_context_data(encrypted1.parms_id_);
            const std::size_t s1 = encrypted1.size_, s2 = encrypted2.size_, so = s1 + s2 - 1;
            if (so > SEAL_CIPHERTEXT_SIZE_MAX)
                throw std::invalid_argument("invalid size");
            const double new_scale = encrypted1.scale_ * encrypted2.scale_;
            if (ckks && !scale_within_bounds(new_scale, *cd))
                throw std::invalid_argument("scale out of bounds");
            std::lock_guard<std::mutex> lock(mu_);
            std::uint64_t *out = out_slab(encrypted1.batch_ * so * encrypted1.L_ * encrypted1.n_, encrypted1);
            check(sb200_multiply_sized(ctx_, encrypted1.L_, s1, s2, encrypted1.batch_, encrypted1.d_, encrypted2.d_, out, nullptr));
            adopt(encrypted1);
            encrypted1.size_ = so;
            if (ckks)
                encrypted1.scale_ = new_scale;
            if (bgv)
                encrypted1.cf_ = seal::util::multiply_uint_mod(encrypted1.cf_, encrypted2.cf_, cd->parms().plain_modulus());
        }
        void square_inplace(CiphertextBatch &encrypted) const { multiply_inplace(encrypted, encrypted); }
        // Evaluator::relinearize_inplace (evaluator.cpp:1144-1199)
        void relinearize_inplace(CiphertextBatch &encrypted, const seal::RelinKeys &relin_keys) const
        {
            owned(encrypted);
            if (relin_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("relin_keys is not valid for encryption parameters");
            if (encrypted.size_ < 2)
                throw std::invalid_argument("destination_size must be at least 2 and less than or equal to current count");
            if (relin_keys.size() < encrypted.size_ - 2)
                throw std::invalid_argument("not enough relinearization keys");
            if (encrypted.size_ == 2)
                return;
            check_keyswitch_form(encrypted.ntt_);
            std::lock_guard<std::mutex> lock(mu_);
            const std::size_t size = encrypted.size_, poly = encrypted.L_ * encrypted.n_;
            if (size == 3)
            {
                sb200_kswitch_key *key = key_for(relin_keys, seal::RelinKeys::get_index(2), encrypted.L_);
                std::uint64_t *out = out_slab(encrypted.batch_ * 2 * poly, encrypted);
                check(sb200_relinearize(ctx_, encrypted.L_, encrypted.batch_, encrypted.d_, key, out, nullptr));
                adopt(encrypted);
                encrypted.size_ = 2;
                return;
            }
            // size > 3, as the reference's loop is written (evaluator.cpp:1176-1187): every step key-switches the LAST polynomial
            // with the key of index size-1-I; afterwards the ciphertexts are cut to 2 polynomials
            for (std::size_t I = 0; I < size - 2; I++)
            {
                sb200_kswitch_key *key = key_for(relin_keys, seal::RelinKeys::get_index(size - 1 - I), encrypted.L_);
                std::uint64_t *out = out_slab(encrypted.batch_ * size * poly, encrypted);
                check(sb200_relinearize_sized(ctx_, encrypted.L_, size, encrypted.batch_, encrypted.d_, key, out, nullptr));
                adopt(encrypted);
            }
            std::uint64_t *out = out_slab(encrypted.batch_ * 2 * poly, encrypted);
            check(sb200_memcpy_d2d_2d(ctx_, out, 2 * poly * sizeof(std::uint64_t), encrypted.d_, size * poly * sizeof(std::uint64_t),
                                      2 * poly * sizeof(std::uint64_t), encrypted.batch_, nullptr));
            adopt(encrypted);
            encrypted.size_ = 2;
        }
        // multiply + relinearize in one device pass (both operands size 2); the result replaces encrypted1
        void multiply_relinearize_inplace(CiphertextBatch &encrypted1, const CiphertextBatch &encrypted2, const seal::RelinKeys &relin_keys) const
        {
            owned(encrypted1), owned(encrypted2);
            if (encrypted1.parms_id_ != encrypted2.parms_id_ || encrypted1.batch_ != encrypted2.batch_)
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            if (relin_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("relin_keys is not valid for encryption parameters");
            if (encrypted1.size_ != 2 || encrypted2.size_ != 2)
            {
                multiply_inplace(encrypted1, encrypted2);
                relinearize_inplace(encrypted1, relin_keys);
                return;
            }
            const bool ckks = scheme_ == seal::scheme_type::ckks, bgv = scheme_ == seal::scheme_type::bgv;
            if ((ckks || bgv) != encrypted1.ntt_ || (ckks || bgv) != encrypted2.ntt_)
                throw std::invalid_argument((ckks || bgv) ? "encrypted1 or encrypted2 must be in NTT form" : "encrypted1 or encrypted2 cannot be in NTT form");
            auto cd = context_.get_context_data(encrypted1.parms_id_);
            const double new_scale = encrypted1.scale_ * encrypted2.scale_;
            if (ckks && !scale_within_bounds(new_scale, *cd))
                throw std::invalid_argument("scale out of bounds");
            if (!context_.using_keyswitching())
                throw std::logic_error("keyswitching is not supported by the context");
            std::lock_guard<std::mutex> lock(mu_);
            sb200_kswitch_key *key = key_for(relin_keys, seal::RelinKeys::get_index(2), encrypted1.L_);
            check(sb200_multiply_relinearize(ctx_, encrypted1.L_, encrypted1.batch_, encrypted1.d_, encrypted2.d_, key, encrypted1.d_, nullptr));
            if (ckks)
                encrypted1.scale_ = new_scale;
            if (bgv)
                encrypted1.cf_ = seal::util::multiply_uint_mod(encrypted1.cf_, encrypted2.cf_, cd->parms().plain_modulus());
        }
        // Evaluator::rescale_to_next_inplace / mod_switch_to_next_inplace (evaluator.cpp:1404-1541), any size
        void rescale_to_next_inplace(CiphertextBatch &encrypted) const
        {
            owned(encrypted);
            if (context_.last_parms_id() == encrypted.parms_id_)
                throw std::invalid_argument("end of modulus switching chain reached");
            if (scheme_ != seal::scheme_type::ckks)
                throw std::invalid_argument("unsupported operation for scheme type");
            mod_switch_batch(encrypted, true);
        }
        void mod_switch_to_next_inplace(CiphertextBatch &encrypted) const
        {
            owned(encrypted);
            if (context_.last_parms_id() == encrypted.parms_id_)
                throw std::invalid_argument("end of modulus switching chain reached");
            mod_switch_batch(encrypted, false);
        }
        // add / sub / negate (evaluator.cpp:130-350) on batches of equal shape
        void add_inplace(CiphertextBatch &encrypted1, const CiphertextBatch &encrypted2) const { linear_batch(encrypted1, encrypted2, false); }
        void sub_inplace(CiphertextBatch &encrypted1, const CiphertextBatch &encrypted2) const { linear_batch(encrypted1, encrypted2, true); }
        void negate_inplace(CiphertextBatch &encrypted) const
        {
            owned(encrypted);
            std::lock_guard<std::mutex> lock(mu_);
            check(sb200_negate(ctx_, encrypted.L_, encrypted.size_, encrypted.batch_, encrypted.d_, encrypted.d_, nullptr));
        }
        // apply_galois / rotations (evaluator.cpp:2384-2559)
        void apply_galois_inplace(CiphertextBatch &encrypted, std::uint32_t galois_elt, const seal::GaloisKeys &galois_keys) const
        {
            owned(encrypted);
            if (galois_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("galois_keys is not valid for encryption parameters");
            if (!galois_keys.has_key(galois_elt))
                throw std::invalid_argument("Galois key not present");
            if (!(galois_elt & 1) || galois_elt >= 2 * encrypted.n_)
                throw std::invalid_argument("Galois element is not valid");
            if (encrypted.size_ != 2)
                throw std::invalid_argument("encrypted size must be 2");
            check_keyswitch_form(encrypted.ntt_);
            std::lock_guard<std::mutex> lock(mu_);
            sb200_kswitch_key *key = key_for(galois_keys, seal::GaloisKeys::get_index(galois_elt), encrypted.L_);
            std::uint64_t *out = out_slab(encrypted.batch_ * encrypted.words_per_ciphertext(), encrypted);
            check(sb200_apply_galois(ctx_, encrypted.L_, encrypted.batch_, encrypted.d_, galois_elt, key, out, nullptr));
            adopt(encrypted);
        }
        void rotate_rows_inplace(CiphertextBatch &encrypted, int steps, const seal::GaloisKeys &galois_keys) const
        {
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::bgv)
                throw std::logic_error("unsupported scheme");
            rotate_batch(encrypted, steps, galois_keys);
        }
        void rotate_vector_inplace(CiphertextBatch &encrypted, int steps, const seal::GaloisKeys &galois_keys) const
        {
            if (scheme_ != seal::scheme_type::ckks)
                throw std::logic_error("unsupported scheme");
            rotate_batch(encrypted, steps, galois_keys);
        }
        void rotate_columns_inplace(CiphertextBatch &encrypted, const seal::GaloisKeys &galois_keys) const
        {
            if (scheme_ != seal::scheme_type::bfv && scheme_ != seal::scheme_type::bgv)
                throw std::logic_error("unsupported scheme");
            conjugate_batch(encrypted, galois_keys);
        }
        void complex_conjugate_inplace(CiphertextBatch &encrypted, const seal::GaloisKeys &galois_keys) const
        {
            if (scheme_ != seal::scheme_type::ckks)
                throw std::logic_error("unsupported scheme");
            conjugate_batch(encrypted, galois_keys);
        }
        // transform_to_ntt_inplace / transform_from_ntt_inplace (evaluator.cpp:2289-2382)
        void transform_to_ntt_inplace(CiphertextBatch &encrypted) const
        {
            owned(encrypted);
            if (encrypted.ntt_)
                throw std::invalid_argument("encrypted is already in NTT form");
            std::lock_guard<std::mutex> lock(mu_);
            check(sb200_ntt_forward(ctx_, encrypted.L_, encrypted.size_, encrypted.batch_, encrypted.d_, nullptr));
            encrypted.ntt_ = true;
        }
        void transform_from_ntt_inplace(CiphertextBatch &encrypted_ntt) const
        {
            owned(encrypted_ntt);
            if (!encrypted_ntt.ntt_)
                throw std::invalid_argument("encrypted_ntt is not in NTT form");
            std::lock_guard<std::mutex> lock(mu_);
            check(sb200_ntt_inverse(ctx_, encrypted_ntt.L_, encrypted_ntt.size_, encrypted_ntt.batch_, encrypted_ntt.d_, nullptr));
            encrypted_ntt.ntt_ = false;
        }
        // waits for all device work queued by this Evaluator (the batch members above only enqueue)
        void synchronize() const { check(sb200_stream_synchronize(ctx_, nullptr)); }

        // ---- key cache control: uploaded key-switching keys are cached by CONTENT (a fingerprint of the key words), never by
        // address; least recently used entries are evicted above the byte budget ----
        void set_key_cache_limit(std::size_t bytes) const
        {
            std::lock_guard<std::mutex> lock(mu_);
            key_cache_limit_ = bytes;
            evict(nullptr);
        }
        void clear_key_cache() const
        {
            std::lock_guard<std::mutex> lock(mu_);
            for (auto &kv : keys_)
                sb200_kswitch_key_destroy(kv.second.handle);
            keys_.clear();
            key_cache_bytes_ = 0;
        }
        std::size_t key_cache_entries() const
        {
            std::lock_guard<std::mutex> lock(mu_);
            return keys_.size();
        }

        sb200_context *native_handle() const noexcept { return ctx_; }

    private:
        // evaluator.cpp:154-262 / 264-350
        void linear(seal::Ciphertext &encrypted1, const seal::Ciphertext &encrypted2, bool subtract) const
        {
            validate(encrypted1, "encrypted1 is not valid for encryption parameters");
            validate(encrypted2, "encrypted2 is not valid for encryption parameters");
            if (encrypted1.parms_id() != encrypted2.parms_id())
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            if (encrypted1.is_ntt_form() != encrypted2.is_ntt_form())
                throw std::invalid_argument("NTT form mismatch");
            if (!seal::util::are_close<double>(encrypted1.scale(), encrypted2.scale()))
                throw std::invalid_argument("scale mismatch");
            auto cd = context_.get_context_data(encrypted1.parms_id());
            const std::size_t L = encrypted1.coeff_modulus_size(), n = encrypted1.poly_modulus_degree();
            if (encrypted1.correction_factor() != encrypted2.correction_factor())
            {
                // BGV: bring both operands to a common correction factor first (evaluator.cpp:188-209, :305-326)
                std::uint64_t f, e1, e2;
                balance_factors(encrypted1.correction_factor(), encrypted2.correction_factor(), cd->parms().plain_modulus(), f, e1, e2);
                seal::Ciphertext copy = encrypted2;
                scale_by(encrypted1, e1, *cd);
                scale_by(copy, e2, *cd);
                encrypted1.correction_factor() = f;
                copy.correction_factor() = f;
                linear(encrypted1, copy, subtract);
                return;
            }
            const std::size_t s1 = encrypted1.size(), s2 = encrypted2.size(), lo = std::min(s1, s2);
            encrypted1.resize(context_, cd->parms_id(), std::max(s1, s2));
            check(subtract ? sb200_sub_host(ctx_, L, lo, 1, encrypted1.data(), encrypted2.data(), encrypted1.data())
                           : sb200_add_host(ctx_, L, lo, 1, encrypted1.data(), encrypted2.data(), encrypted1.data()));
            if (s1 < s2)
            {
                // the longer operand's tail is copied (add, :228-233) or negated (sub, :336-341)
                if (subtract)
                    check(sb200_negate_host(ctx_, L, s2 - lo, 1, encrypted2.data(lo), encrypted1.data(lo)));
                else
                    std::memcpy(encrypted1.data(lo), encrypted2.data(lo), (s2 - lo) * L * n * sizeof(std::uint64_t));
            }
            throw_if_transparent(encrypted1);
        }
        // every polynomial times the scalar e (multiply_poly_scalar_coeffmod, evaluator.cpp:192-200): a dyadic product with the
        // constant vector e mod q_i, valid in either form
        void scale_by(seal::Ciphertext &ct, std::uint64_t e, const seal::SEALContext::ContextData &cd) const
        {
            const std::size_t L = ct.coeff_modulus_size(), n = ct.poly_modulus_degree();
            std::vector<std::uint64_t> v(L * n);
            for (std::size_t i = 0; i < L; i++)
                std::fill(v.begin() + i * n, v.begin() + (i + 1) * n, seal::util::barrett_reduce_64(e, cd.parms().coeff_modulus()[i]));
            check(sb200_multiply_plain_host(ctx_, L, ct.size(), 1, ct.data(), v.data(), ct.data()));
        }
        // (f, e1, e2) with e1*factor1 = e2*factor2 = f mod t, both e invertible mod t, and |e1| + |e2| (balanced
        // representatives) minimal over the remainder sequence of (t, factor2/factor1) -- evaluator.cpp:50-118
        static void balance_factors(std::uint64_t factor1, std::uint64_t factor2, const seal::Modulus &plain, std::uint64_t &f,
                                    std::uint64_t &e1, std::uint64_t &e2)
        {
            using namespace seal::util;
            const std::uint64_t t = plain.value();
            auto weight = [t](std::uint64_t x, std::uint64_t y) {
                auto mag = [t](std::uint64_t v) { return static_cast<std::int64_t>(v > t / 2 ? t - v : v); };
                return mag(x) + mag(y);
            };
            auto residue = [&plain](std::int64_t v) {
                std::uint64_t r = barrett_reduce_64(static_cast<std::uint64_t>(v < 0 ? -v : v), plain);
                return v < 0 ? negate_uint_mod(r, plain) : r;
            };
            std::uint64_t ratio = 1;
            if (!try_invert_uint_mod(factor1, plain, ratio))
                throw std::logic_error("invalid correction factor1");
            ratio = multiply_uint_mod(ratio, factor2, plain);
            e1 = ratio, e2 = 1;
            std::int64_t best = weight(e1, e2);
            std::int64_t r0 = static_cast<std::int64_t>(t), r1 = static_cast<std::int64_t>(ratio), c0 = 0, c1 = 1;
            while (r1 != 0)
            {
                const std::int64_t quo = r0 / r1, r2 = r0 % r1, c2 = sub_safe(c0, mul_safe(c1, quo));
                r0 = r1, r1 = r2, c0 = c1, c1 = c2;
                const std::uint64_t a = residue(r1), b = residue(c1);
                if (a != 0 && gcd(a, t) == 1)
                {
                    const std::int64_t w = weight(a, b);
                    if (w < best)
                        best = w, e1 = a, e2 = b;
                }
            }
            f = multiply_uint_mod(e1, factor1, plain);
        }
        // a coefficient-form plaintext as n words (Plaintext::coeff_count() may be smaller; the rest is zero)
        std::vector<std::uint64_t> padded(const seal::Plaintext &plain) const
        {
            const std::size_t n = context_.key_context_data()->parms().poly_modulus_degree();
            std::vector<std::uint64_t> w(n, 0);
            std::copy(plain.data(), plain.data() + std::min(n, plain.coeff_count()), w.begin());
            return w;
        }
        // mod_switch_drop_to_next (evaluator.cpp:1296-1366): every polynomial keeps its first L-1 components
        void drop_last_component(seal::Ciphertext &encrypted) const
        {
            auto cd = context_.get_context_data(encrypted.parms_id());
            const auto scheme = cd->parms().scheme();
            if (scheme == seal::scheme_type::bfv && encrypted.is_ntt_form())
                throw std::invalid_argument("BFV encrypted cannot be in NTT form");
            if (scheme == seal::scheme_type::ckks && !encrypted.is_ntt_form())
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (scheme == seal::scheme_type::bgv && !encrypted.is_ntt_form())
                throw std::invalid_argument("BGV encrypted must be in NTT form");
            auto &next = *cd->next_context_data();
            if (!scale_within_bounds(encrypted.scale(), next))
                throw std::invalid_argument("scale out of bounds");
            const std::size_t L = encrypted.coeff_modulus_size(), n = encrypted.poly_modulus_degree(), size = encrypted.size();
            std::vector<std::uint64_t> kept(size * (L - 1) * n);
            for (std::size_t p = 0; p < size; p++)
                std::memcpy(kept.data() + p * (L - 1) * n, encrypted.data(p), (L - 1) * n * sizeof(std::uint64_t));
            const bool ntt = encrypted.is_ntt_form();
            const double scale = encrypted.scale();
            const std::uint64_t correction = encrypted.correction_factor();
            encrypted.resize(context_, next.parms_id(), size);
            std::memcpy(encrypted.data(), kept.data(), kept.size() * sizeof(std::uint64_t));
            encrypted.is_ntt_form() = ntt, encrypted.scale() = scale, encrypted.correction_factor() = correction;
        }
        void check_target_level(const seal::Ciphertext &encrypted, const seal::parms_id_type &parms_id) const
        {
            auto cur = context_.get_context_data(encrypted.parms_id());
            auto target = context_.get_context_data(parms_id);
            if (!cur)
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (!target)
                throw std::invalid_argument("parms_id is not valid for encryption parameters");
            if (cur->chain_index() < target->chain_index())
                throw std::invalid_argument("cannot switch to higher level modulus");
        }
        void plain_linear(seal::Ciphertext &encrypted, const seal::Plaintext &plain, bool subtract) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (!seal::is_metadata_valid_for(plain, context_) || !seal::is_buffer_valid(plain))
                throw std::invalid_argument("plain is not valid for encryption parameters");
            if (scheme_ != seal::scheme_type::ckks)
            {
                // BFV: scaling variant on c_0 (util/scalingvariant.cpp:70-160); BGV: correction factor, lift, NTT (:1838-1849)
                const bool bfv = scheme_ == seal::scheme_type::bfv;
                if (bfv && encrypted.is_ntt_form())
                    throw std::invalid_argument("BFV encrypted cannot be in NTT form");
                if (!bfv && !encrypted.is_ntt_form())
                    throw std::invalid_argument("BGV encrypted must be in NTT form");
                if (plain.is_ntt_form())
                    throw std::invalid_argument(bfv ? "BFV plain cannot be in NTT form" : "BGV plain cannot be in NTT form");
                const std::vector<std::uint64_t> words = padded(plain);
                const std::uint64_t cf = encrypted.correction_factor();
                check(sb200_add_plain_coeff_host(ctx_, encrypted.coeff_modulus_size(), encrypted.size(), 1, subtract ? 1 : 0, encrypted.data(),
                                                 words.data(), bfv ? nullptr : &cf, encrypted.data()));
                throw_if_transparent(encrypted);
                return;
            }
            if (!encrypted.is_ntt_form())
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (!plain.is_ntt_form())
                throw std::invalid_argument("CKKS plain must be in NTT form");
            if (encrypted.parms_id() != plain.parms_id())
                throw std::invalid_argument("encrypted and plain parameter mismatch");
            if (!seal::util::are_close<double>(encrypted.scale(), plain.scale()))
                throw std::invalid_argument("scale mismatch");
            const std::size_t L = encrypted.coeff_modulus_size();
            // c_0 +/- plain: a one-polynomial slab (add_poly_coeffmod / sub_poly_coeffmod on encrypted[0], :1831-1836)
            check(subtract ? sb200_sub_host(ctx_, L, 1, 1, encrypted.data(), plain.data(), encrypted.data())
                           : sb200_add_host(ctx_, L, 1, 1, encrypted.data(), plain.data(), encrypted.data()));
            throw_if_transparent(encrypted);
        }
        static void check(int status)
        {
            if (status == SB200_OK)
                return;
            const std::string msg = sb200_last_error();
            switch (status)
            {
            case SB200_E_INVALID_ARG:
            case SB200_E_POINTER: throw std::invalid_argument(msg);
            case SB200_E_LOGIC: throw std::logic_error(msg);
            case SB200_E_OUT_OF_RANGE: throw std::out_of_range(msg);
            case SB200_E_NOMEM: throw std::bad_alloc();
            default: throw std::runtime_error(msg);
            }
        }
        void validate(const seal::Ciphertext &ct, const char *msg) const
        {
            if (!seal::is_metadata_valid_for(ct, context_) || !seal::is_buffer_valid(ct))
                throw std::invalid_argument(msg);
        }
        static void throw_if_transparent(const seal::Ciphertext &ct)
        {
#ifdef SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT
            if (ct.is_transparent())
                throw std::logic_error("result ciphertext is transparent");
#else
            (void)ct;
#endif
        }
        // evaluator.cpp:29-48
        static bool scale_within_bounds(double scale, const seal::SEALContext::ContextData &cd) noexcept
        {
            int bound = cd.parms().scheme() == seal::scheme_type::ckks ? cd.total_coeff_modulus_bit_count()
                                                                        : cd.parms().plain_modulus().bit_count();
            return !(!std::isnormal(scale) || scale <= 0 || (static_cast<int>(std::log2(scale)) >= bound));
        }
        // the parts of switch_key_inplace's prologue that concern the operand (evaluator.cpp:2573-2611)
        void check_keyswitch_operand(const seal::Ciphertext &encrypted, const seal::MemoryPoolHandle &pool) const
        {
            validate(encrypted, "encrypted is not valid for encryption parameters");
            if (!context_.using_keyswitching())
                throw std::logic_error("keyswitching is not supported by the context");
            if (!pool)
                throw std::invalid_argument("pool is uninitialized");
            if (scheme_ == seal::scheme_type::bfv && encrypted.is_ntt_form())
                throw std::invalid_argument("BFV encrypted cannot be in NTT form");
            if (scheme_ == seal::scheme_type::ckks && !encrypted.is_ntt_form())
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (scheme_ == seal::scheme_type::bgv && !encrypted.is_ntt_form())
                throw std::invalid_argument("BGV encrypted must be in NTT form");
        }
        // uploads KSwitchKeys::data()[index] (evaluator.cpp:2586-2648) and caches the device copy under a fingerprint of the
        // key's CONTENT (leading / trailing / strided words of every digit + its shape): a regenerated or reloaded key object
        // that happens to reuse a freed pool address can never be served another key's device copy.  Caller holds mu_.
        struct KeyEntry
        {
            sb200_kswitch_key *handle = nullptr;
            std::size_t bytes = 0;
            std::uint64_t last_use = 0;
        };
        static std::uint64_t fingerprint(const std::vector<seal::PublicKey> &kv, std::size_t row_words, std::size_t index)
        {
            std::uint64_t h = 0x9E3779B97F4A7C15ull ^ (row_words * 0xD6E8FEB86659FD93ull) ^ (kv.size() << 48) ^ (index << 32);
            auto mix = [&h](std::uint64_t v) {
                h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
                h *= 0xFF51AFD7ED558CCDull;
                h ^= h >> 33;
            };
            const std::size_t edge = std::min<std::size_t>(64, row_words), stride = std::max<std::size_t>(1, row_words / 96) | 1;
            for (auto &pk : kv)
            {
                const std::uint64_t *p = pk.data().data();
                for (std::size_t i = 0; i < edge; i++)
                    mix(p[i]), mix(p[row_words - 1 - i]);
                for (std::size_t i = 0; i < row_words; i += stride)
                    mix(p[i]);
            }
            return h;
        }
        void evict(const sb200_kswitch_key *keep) const
        {
            while (key_cache_bytes_ > key_cache_limit_ && keys_.size() > (keep ? 1u : 0u))
            {
                auto victim = keys_.end();
                for (auto it = keys_.begin(); it != keys_.end(); ++it)
                    if (it->second.handle != keep && (victim == keys_.end() || it->second.last_use < victim->second.last_use))
                        victim = it;
                if (victim == keys_.end())
                    break;
                sb200_kswitch_key_destroy(victim->second.handle);
                key_cache_bytes_ -= victim->second.bytes;
                keys_.erase(victim);
            }
        }
        sb200_kswitch_key *key_for(const seal::KSwitchKeys &keys, std::size_t index, std::size_t L) const
        {
            if (keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("parameter mismatch");
            if (index >= keys.data().size())
                throw std::out_of_range("kswitch_keys_index");
            auto &kv = keys.data(index); // throws on an empty slot, like the reference's checked accessor
            if (kv.size() < L)
                throw std::invalid_argument("kswitch_keys inner dimension is too small");
            for (auto &each : kv)
                if (!seal::is_metadata_valid_for(each, context_) || !seal::is_buffer_valid(each))
                    throw std::invalid_argument("kswitch_keys is not valid for encryption parameters");
            const std::size_t K = context_.key_context_data()->parms().coeff_modulus().size();
            const std::size_t n = context_.key_context_data()->parms().poly_modulus_degree(), row = 2 * K * n;
            const std::uint64_t id = fingerprint(kv, row, index);
            auto it = keys_.find(id);
            if (it != keys_.end())
            {
                it->second.last_use = ++key_clock_;
                return it->second.handle;
            }
            std::vector<std::uint64_t> flat(kv.size() * row);
            for (std::size_t j = 0; j < kv.size(); j++)
                std::memcpy(flat.data() + j * row, kv[j].data().data(), row * sizeof(std::uint64_t));
            sb200_kswitch_key *h = nullptr;
            check(sb200_kswitch_key_create(ctx_, flat.data(), kv.size(), &h));
            KeyEntry e;
            e.handle = h, e.bytes = flat.size() * sizeof(std::uint64_t), e.last_use = ++key_clock_;
            keys_[id] = e;
            key_cache_bytes_ += e.bytes;
            evict(h);
            return h;
        }
        // ---- helpers of the CiphertextBatch members ----
        // data of already validated ciphertexts of one shape -> one device slab; metadata of the first member
        void upload_rows(const seal::Ciphertext *cts, std::size_t count, CiphertextBatch &destination) const
        {
            const seal::Ciphertext &f = cts[0];
            {
                std::lock_guard<std::mutex> lock(mu_);
                shape(destination, count, f.size(), f.coeff_modulus_size(), f.poly_modulus_degree());
            }
            destination.parms_id_ = f.parms_id(), destination.ntt_ = f.is_ntt_form(), destination.scale_ = f.scale();
            destination.cf_ = f.correction_factor();
            // the transfer itself runs outside the lock (its own staging lane and stream in the library): uploads of one batch
            // overlap the operations another thread enqueues on other batches
            std::vector<const std::uint64_t *> rows(count);
            for (std::size_t i = 0; i < count; i++)
                rows[i] = cts[i].data();
            check(sb200_upload_rows(ctx_, destination.d_, rows.data(), destination.words_per_ciphertext() * sizeof(std::uint64_t), count));
        }
        void owned(const CiphertextBatch &b) const
        {
            if (b.ctx_ != ctx_ || !b.d_ || !b.batch_)
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (!context_.get_context_data(b.parms_id_))
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
        }
        void reserve(CiphertextBatch &b, std::size_t words) const
        {
            if (b.ctx_ != ctx_)
            {
                b.release();
                b.ctx_ = ctx_;
            }
            if (b.cap_ < words)
            {
                b.release();
                check(sb200_device_malloc(ctx_, words * sizeof(std::uint64_t), &b.d_));
                b.cap_ = words;
            }
        }
        void shape(CiphertextBatch &b, std::size_t batch, std::size_t size, std::size_t L, std::size_t n) const
        {
            reserve(b, batch * size * L * n);
            b.batch_ = batch, b.size_ = size, b.L_ = L, b.n_ = n;
        }
        static void copy_meta(const CiphertextBatch &s, CiphertextBatch &d)
        {
            d.parms_id_ = s.parms_id_, d.ntt_ = s.ntt_, d.scale_ = s.scale_, d.cf_ = s.cf_;
        }
        // result slab of a layout-changing operation; adopt() swaps it with the operand's storage, so a chain of operations
        // allocates nothing once the largest shape has been seen.  Caller holds mu_.
        // (never smaller than the slab it will be swapped with: every buffer in circulation then fits every level of a chain,
        // and a later upload into the same batch object finds room without reallocating)
        std::uint64_t *out_slab(std::size_t words, const CiphertextBatch &operand) const
        {
            reserve(tmp_, std::max(words, operand.cap_));
            return tmp_.d_;
        }
        void adopt(CiphertextBatch &b) const
        {
            std::swap(b.d_, tmp_.d_);
            std::swap(b.cap_, tmp_.cap_);
        }
        void check_keyswitch_form(bool ntt) const
        {
            if (!context_.using_keyswitching())
                throw std::logic_error("keyswitching is not supported by the context");
            if (scheme_ == seal::scheme_type::bfv && ntt)
                throw std::invalid_argument("BFV encrypted cannot be in NTT form");
            if (scheme_ == seal::scheme_type::ckks && !ntt)
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (scheme_ == seal::scheme_type::bgv && !ntt)
                throw std::invalid_argument("BGV encrypted must be in NTT form");
        }
        void mod_switch_batch(CiphertextBatch &e, bool rescale) const
        {
            auto cd = context_.get_context_data(e.parms_id_);
            auto next = cd->next_context_data();
            const bool ckks = scheme_ == seal::scheme_type::ckks, bgv = scheme_ == seal::scheme_type::bgv;
            if (ckks && !e.ntt_)
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (bgv && !e.ntt_)
                throw std::invalid_argument("BGV encrypted must be in NTT form");
            if (!ckks && !bgv && e.ntt_)
                throw std::invalid_argument("BFV encrypted cannot be in NTT form");
            double scale = e.scale_;
            if (rescale)
            {
                if (!scale_within_bounds(e.scale_, *cd))
                    throw std::invalid_argument("scale out of bounds");
                scale = e.scale_ / static_cast<double>(cd->parms().coeff_modulus().back().value());
            }
            if ((rescale || ckks) && !scale_within_bounds(scale, *next))
                throw std::invalid_argument("scale out of bounds"); // :1236-1241, mod_switch_drop_to_next :1318-1322
            std::lock_guard<std::mutex> lock(mu_);
            std::uint64_t *out = out_slab(e.batch_ * e.size_ * (e.L_ - 1) * e.n_, e);
            check(rescale ? sb200_rescale_to_next_sized(ctx_, e.L_, e.size_, e.batch_, e.d_, out, nullptr)
                          : sb200_mod_switch_to_next_sized(ctx_, e.L_, e.size_, e.batch_, e.d_, out, nullptr));
            adopt(e);
            e.L_ -= 1;
            e.parms_id_ = next->parms_id();
            e.scale_ = scale;
            if (bgv)
                e.cf_ = seal::util::multiply_uint_mod(e.cf_, cd->rns_tool()->inv_q_last_mod_t(), next->parms().plain_modulus());
        }
        void linear_batch(CiphertextBatch &a, const CiphertextBatch &b, bool subtract) const
        {
            owned(a), owned(b);
            if (a.parms_id_ != b.parms_id_ || a.batch_ != b.batch_)
                throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
            if (a.ntt_ != b.ntt_)
                throw std::invalid_argument("NTT form mismatch");
            if (!seal::util::are_close<double>(a.scale_, b.scale_))
                throw std::invalid_argument("scale mismatch");
            if (a.size_ != b.size_ || a.cf_ != b.cf_)
                throw std::logic_error("seal_b200: batch add / sub needs operands of equal size and correction factor");
            std::lock_guard<std::mutex> lock(mu_);
            check(subtract ? sb200_sub(ctx_, a.L_, a.size_, a.batch_, a.d_, b.d_, a.d_, nullptr)
                           : sb200_add(ctx_, a.L_, a.size_, a.batch_, a.d_, b.d_, a.d_, nullptr));
        }
        void rotate_batch(CiphertextBatch &encrypted, int steps, const seal::GaloisKeys &galois_keys) const
        {
            owned(encrypted);
            auto cd = context_.get_context_data(encrypted.parms_id_);
            if (!cd->qualifiers().using_batching)
                throw std::logic_error("encryption parameters do not support batching");
            if (galois_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("galois_keys is not valid for encryption parameters");
            if (steps == 0)
                return;
            const std::size_t n = cd->parms().poly_modulus_degree();
            const std::uint32_t elt = cd->galois_tool()->get_elt_from_step(steps);
            if (galois_keys.has_key(elt))
            {
                apply_galois_inplace(encrypted, elt, galois_keys);
                return;
            }
            std::vector<int> naf_steps = seal::util::naf(steps);
            if (naf_steps.size() == 1)
                throw std::invalid_argument("Galois key not present");
            for (int st : naf_steps)
                if (static_cast<std::size_t>(std::abs(st)) != (n >> 1))
                    rotate_batch(encrypted, st, galois_keys);
        }
        void conjugate_batch(CiphertextBatch &encrypted, const seal::GaloisKeys &galois_keys) const
        {
            owned(encrypted);
            auto cd = context_.get_context_data(encrypted.parms_id_);
            if (!cd->qualifiers().using_batching)
                throw std::logic_error("encryption parameters do not support batching");
            apply_galois_inplace(encrypted, cd->galois_tool()->get_elt_from_step(0), galois_keys);
        }
        void mod_switch_impl(const seal::Ciphertext &encrypted, seal::Ciphertext &destination, bool rescale) const
        {
            auto cd = context_.get_context_data(encrypted.parms_id());
            auto next = cd->next_context_data();
            const bool ckks = scheme_ == seal::scheme_type::ckks, bgv = scheme_ == seal::scheme_type::bgv;
            if (ckks && !encrypted.is_ntt_form())
                throw std::invalid_argument("CKKS encrypted must be in NTT form");
            if (bgv && !encrypted.is_ntt_form())
                throw std::invalid_argument("BGV encrypted must be in NTT form"); // :1214-1217
            if (!ckks && !bgv && encrypted.is_ntt_form())
                throw std::invalid_argument("BFV encrypted cannot be in NTT form");
            double scale = encrypted.scale();
            if (rescale)
            {
                if (!scale_within_bounds(encrypted.scale(), *cd))
                    throw std::invalid_argument("scale out of bounds");
                scale = encrypted.scale() / static_cast<double>(cd->parms().coeff_modulus().back().value());
                if (!scale_within_bounds(scale, *next))
                    throw std::invalid_argument("scale out of bounds");
            }
            else if (ckks && !scale_within_bounds(scale, *next))
                throw std::invalid_argument("scale out of bounds"); // mod_switch_drop_to_next, evaluator.cpp:1318-1322
            const std::size_t L = encrypted.coeff_modulus_size(), n = encrypted.poly_modulus_degree(), size = encrypted.size();
            std::vector<std::uint64_t> in(encrypted.data(), encrypted.data() + size * L * n);
            const bool ntt = encrypted.is_ntt_form();
            const std::uint64_t correction = encrypted.correction_factor();
            destination.resize(context_, next->parms_id(), size);
            check(rescale ? sb200_rescale_to_next_sized_host(ctx_, L, size, 1, in.data(), destination.data())
                          : sb200_mod_switch_to_next_sized_host(ctx_, L, size, 1, in.data(), destination.data()));
            destination.is_ntt_form() = ntt;
            destination.scale() = scale;
            if (bgv) // :1288-1293
                destination.correction_factor() =
                    seal::util::multiply_uint_mod(correction, cd->rns_tool()->inv_q_last_mod_t(), next->parms().plain_modulus());
            throw_if_transparent(destination);
        }
        // evaluator.cpp:2504-2559
        void rotate_internal(seal::Ciphertext &encrypted, int steps, const seal::GaloisKeys &galois_keys, seal::MemoryPoolHandle pool) const
        {
            auto cd = context_.get_context_data(encrypted.parms_id());
            if (!cd)
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (!cd->qualifiers().using_batching)
                throw std::logic_error("encryption parameters do not support batching");
            if (galois_keys.parms_id() != context_.key_parms_id())
                throw std::invalid_argument("galois_keys is not valid for encryption parameters");
            if (steps == 0)
                return;
            const std::size_t n = cd->parms().poly_modulus_degree();
            const std::uint32_t elt = cd->galois_tool()->get_elt_from_step(steps);
            if (galois_keys.has_key(elt))
            {
                apply_galois_inplace(encrypted, elt, galois_keys, std::move(pool));
                return;
            }
            std::vector<int> naf_steps = seal::util::naf(steps);
            if (naf_steps.size() == 1)
                throw std::invalid_argument("Galois key not present");
            for (int s : naf_steps)
                if (static_cast<std::size_t>(std::abs(s)) != (n >> 1))
                    rotate_internal(encrypted, s, galois_keys, pool);
        }
        void conjugate_internal(seal::Ciphertext &encrypted, const seal::GaloisKeys &galois_keys, seal::MemoryPoolHandle pool) const
        {
            auto cd = context_.get_context_data(encrypted.parms_id());
            if (!cd)
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (!cd->qualifiers().using_batching)
                throw std::logic_error("encryption parameters do not support batching");
            apply_galois_inplace(encrypted, cd->galois_tool()->get_elt_from_step(0), galois_keys, std::move(pool));
        }

        seal::SEALContext context_;
        seal::scheme_type scheme_;
        sb200_context *ctx_ = nullptr;
        mutable std::mutex mu_;
        mutable std::map<std::uint64_t, KeyEntry> keys_;
        mutable std::size_t key_cache_bytes_ = 0, key_cache_limit_ = std::size_t(24) << 30;
        mutable std::uint64_t key_clock_ = 0;
        mutable CiphertextBatch tmp_; // result slab of layout-changing batch operations
    };
} // namespace seal_b200
