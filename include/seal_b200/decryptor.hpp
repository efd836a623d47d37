// include/seal_b200/decryptor.hpp -- same-signature stand-in for seal::Decryptor::decrypt (native/src/seal/decryptor.h)
// over the C-ABI of include/seal_b200.h: the dot product with the secret key, BFV's scale-and-round, BGV's exact base
// conversion run on the device (sb200_decrypt); checks, exception types and the shape of the resulting Plaintext are the
// reference's (decryptor.cpp:53-197).  invariant_noise_budget stays with seal::Decryptor.
#pragma once

#include "../seal_b200.h"
#include "evaluator.hpp"
#include "seal/seal.h"
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace seal_b200
{
    class Decryptor
    {
    public:
        // shares the device context of an existing evaluator, which must outlive this object
        Decryptor(const seal::SEALContext &context, const seal::SecretKey &secret_key, const Evaluator &evaluator)
            : context_(context), ctx_(evaluator.native_handle())
        {
            if (!context_.parameters_set())
                throw std::invalid_argument("encryption parameters are not set correctly");
            if (!seal::is_valid_for(secret_key, context_))
                throw std::invalid_argument("secret key is not valid for encryption parameters");
            status(sb200_secret_key_create(ctx_, secret_key.data().data(), &key_));
        }
        ~Decryptor()
        {
            if (key_)
                sb200_secret_key_destroy(key_);
        }
        Decryptor(const Decryptor &) = delete;
        Decryptor &operator=(const Decryptor &) = delete;

        void decrypt(const seal::Ciphertext &encrypted, seal::Plaintext &destination)
        {
            if (!seal::is_valid_for(encrypted, context_))
                throw std::invalid_argument("encrypted is not valid for encryption parameters");
            if (encrypted.size() < SEAL_CIPHERTEXT_SIZE_MIN)
                throw std::invalid_argument("encrypted is empty");
            const auto scheme = context_.first_context_data()->parms().scheme();
            const std::size_t L = encrypted.coeff_modulus_size(), n = encrypted.poly_modulus_degree();
            if (scheme == seal::scheme_type::bfv && encrypted.is_ntt_form())
                throw std::invalid_argument("encrypted cannot be in NTT form"); // :113-116
            if (scheme != seal::scheme_type::bfv && !encrypted.is_ntt_form())
                throw std::invalid_argument("encrypted must be in NTT form"); // :139-142, :161-164
            destination.parms_id() = seal::parms_id_zero;
            if (scheme == seal::scheme_type::ckks)
            {
                destination.resize(L * n);
                status(sb200_decrypt_host(ctx_, key_, L, encrypted.size(), 1, encrypted.data(), nullptr, destination.data()));
                destination.parms_id() = encrypted.parms_id();
                destination.scale() = encrypted.scale();
                return;
            }
            destination.resize(n);
            const std::uint64_t cf = encrypted.correction_factor();
            status(sb200_decrypt_host(ctx_, key_, L, encrypted.size(), 1, encrypted.data(), scheme == seal::scheme_type::bgv ? &cf : nullptr,
                                      destination.data()));
            // the plaintext keeps its significant coefficients only (:131-134, :193-196)
            std::size_t count = n;
            while (count > 1 && destination.data()[count - 1] == 0)
                count--;
            destination.resize(count);
        }

    private:
        static void status(int rc)
        {
            if (rc == SB200_OK)
                return;
            const std::string msg = sb200_last_error();
            if (rc == SB200_E_INVALID_ARG || rc == SB200_E_POINTER)
                throw std::invalid_argument(msg);
            if (rc == SB200_E_LOGIC)
                throw std::logic_error(msg);
            throw std::runtime_error(msg);
        }
        seal::SEALContext context_;
        sb200_context *ctx_ = nullptr;
        sb200_secret_key *key_ = nullptr;
    };
} // namespace seal_b200
