// seal_b200/csrc/sb_wire.hpp -- the reference's wire format for ciphertexts, host side only (no CUDA):
//   * parms_id = BLAKE2b-256 over [scheme, n, q_0..q_{L-1}, t]            (encryptionparams.cpp:124-158, util/hash.h:30-38)
//   * Ciphertext::save with compr_mode_type::none = SEALHeader + members + DynArray (ciphertext.cpp:190-247,
//     serialization.h:76-91, dynarray.h:662-690)
// so that serialized ciphertexts move between a byte stream and a device slab without a seal::Ciphertext in between
// (SURVEY 8f rank 3).  Compressed streams and seed-compressed ciphertexts stay with the reference.
#pragma once
#include "../../include/seal_b200.h"
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sbw
{
    using u64 = unsigned long long;

    // unkeyed BLAKE2b with a 32-byte digest (RFC 7693), digest returned as 4 little-endian words
    void blake2b_256(const void *in, size_t len, u64 out[4]);
    // EncryptionParameters::compute_parms_id for the level with primes q[0..L)
    void parms_id(int scheme, size_t n, const u64 *q, size_t L, u64 t, u64 out[4]);

    constexpr size_t kHeaderBytes = 16;                              // Serialization::SEALHeader
    constexpr size_t kMemberBytes = 32 + 1 + 8 + 8 + 8 + 8 + 8;      // parms_id .. correction_factor
    constexpr size_t kDataOffset = kHeaderBytes + kMemberBytes + kHeaderBytes + 8; // first coefficient word

    // Serialization::Load of a zlib-compressed object (serialization.cpp:236-300, util/ztools.cpp): SEALHeader{compr_mode = zlib,
    // size} followed by ONE zlib stream of the object's body.  Returns false for an uncompressed stream; otherwise fills `plain`
    // with the equivalent uncompressed stream (header with compr_mode none + body) and returns true.  `limit` bounds the
    // decompressed body (the reference bounds loads by the expected in-memory size for the same reason: a malformed stream must
    // not cause arbitrarily large allocations).  zstd streams throw (std::logic_error): the library is not part of this build.
    bool inflate_stream(const uint8_t *p, size_t len, size_t limit, std::vector<uint8_t> &plain);
    // parses and validates one serialized ciphertext (throws std::invalid_argument / std::logic_error like Serialization::Load)
    void inspect(const uint8_t *p, size_t len, sb200_ct_info &info);
    inline size_t save_size(size_t words) { return kDataOffset + 8 * words; }
    // writes everything in front of the coefficient words (kDataOffset bytes) for a ciphertext of info.data_words words
    void write_prefix(const sb200_ct_info &info, uint8_t *out);
    // One entry of a serialized KSwitchKeys object (RelinKeys / GaloisKeys saved with compr_mode_type::none,
    // kswitchkeys.cpp:42-86: SEALHeader, parms_id, dim1, then per slot dim2 and dim2 PublicKey = Ciphertext streams).
    // Returns the byte offsets (from p) of the coefficient words of data()[index][j], j < digits, each 2*L*n words.
    struct KSwitchEntry
    {
        u64 parms_id[4];
        size_t slots = 0;               // data().size()
        size_t L = 0, n = 0;            // shape of every key polynomial pair
        std::vector<size_t> offsets;    // one per decomposition digit
    };
    void inspect_kswitch(const uint8_t *p, size_t len, size_t index, KSwitchEntry &entry);
} // namespace sbw
