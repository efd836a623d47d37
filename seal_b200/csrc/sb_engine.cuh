// seal_b200/csrc/sb_engine.cuh -- device context + operation drivers (declarations).
#pragma once
#include "sb_host.hpp"
#include "sb_ntt.cuh"
#include "sb_ksint.cuh"
#include "sb_wire.hpp"
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace sb
{
    struct CudaError : std::runtime_error
    {
        using std::runtime_error::runtime_error;
    };
    void cuda_check(cudaError_t e, const char *what);

    // device copy of sbh::BehzLevel (BFV multiply), see sb_bfv.cu
    struct BehzDev;

    // the secret key on the device for decryption: powers s^1 .. s^powers in NTT form at the key level, [powers][k][n]
    struct SecretKey
    {
        struct Context *ctx = nullptr;
        u64 *d_pow = nullptr;
        size_t powers = 0;
    };

    // PublicKey::data() on the device: [2][k][n], NTT form at the key level
    struct PublicKey
    {
        struct Context *ctx = nullptr;
        u64 *d_key = nullptr;
    };

    struct KSwitchKey
    {
        struct Context *ctx = nullptr;
        u64 *d_key = nullptr; // [digits][2][k][n]
        uint32_t *d_key32 = nullptr; // [S][digits][2][k][n]: the key modulo the auxiliary primes, transformed (sb_ksint.cuh)
        size_t digits = 0;
    };

    // staging for the host-buffer entry points (sb_api.cu: HostPipe)
    struct IoArena
    {
        bool ready = false;
        cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
        cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
        u64 *buf[2][3] = {};
        size_t cap[2][3] = {};
        // page-locked staging of sb200_upload_rows ([0]) / sb200_download_rows ([1]): separate double buffers, streams and locks,
        // so that an upload, a download and the kernels of a third batch overlap (three host threads, or one thread per direction)
        struct Lane
        {
            void *pin[2] = {};
            size_t cap = 0;
            cudaEvent_t ev[2] = {}, order = nullptr;
            cudaStream_t st = nullptr;
            std::mutex mu;
        } lane[2];
    };

    struct Context
    {
        int scheme = 0, device = 0, logn = 0;
        bool fast_q = true;                  // every q prime < 2^57: guard-free forward butterflies (sb_device.cuh)
        size_t n = 0, k = 0;
        u64 t = 0;
        std::vector<u64> q;                  // key-level primes
        std::vector<sbh::PrimeTables> tabs;  // host tables; ids [0,k) = q primes, [k, ...) = BEHZ auxiliary primes
        std::vector<u64> aux;                // auxiliary prime list [m_sk, gamma, B_0, B_1, ...] (BFV)
        std::vector<Tw *> d_fwd, d_inv;      // per prime id
        PrimeDev *d_primes = nullptr;        // [nprimes]
        size_t nprimes = 0;
        Tw *d_invq = nullptr;                // [k][k]: d_invq[j*k+i] = q_j^-1 mod q_i  (i != j)
        Tw *d_qmod = nullptr;                // BGV, [k][k]: d_qmod[j*k+i] = q_j mod q_i
        u64 t_ratio = 0;                     // BGV: floor(2^64 / t)
        std::vector<u64> inv_q_mod_t;        // BGV: q_j^-1 mod t
        // decryption constants per level (sb_engine.cu: decrypt_level): [L] x {Tw c, u64 m_t, u64 m_g}
        struct DecryptLevel
        {
            void *d_consts = nullptr;
            u64 neg_inv_q_mod_t = 0, neg_inv_q_mod_g = 0, q_mod_t = 0;
        };
        std::map<size_t, DecryptLevel> decrypt_levels;
        // BatchEncoder on the device (BFV / BGV with an NTT-friendly plain modulus): prime id of t and the inverse slot map
        int t_pid = -1;
        uint32_t *d_batch_inv_map = nullptr; // coefficient index -> matrix slot
        // coefficient-form plaintext operations (BFV / BGV): per-level constants, built on first use (sb_engine.cu: plain_level)
        struct PlainLevel
        {
            Tw *d_delta = nullptr;  // [L]: floor(q / t) mod q_i (ContextData::coeff_div_plain_modulus, context.cpp:300-318)
            u64 q_mod_t = 0;        // ContextData::coeff_modulus_mod_plain_modulus
        };
        std::map<size_t, PlainLevel> plain_levels;
        // CKKSEncoder on the device (sb_ckks.cu): complex root tables + slot map, built on first use; per-level big integers for decode
        struct CkksEncoder
        {
            bool ready = false;
            void *d_roots = nullptr, *d_inv_roots = nullptr; // double2 [n]
            uint32_t *d_map = nullptr;                       // matrix_reps_index_map_
            void *d_stat = nullptr;                          // { max |coefficient| bits, flags }
        } ckks;
        struct CkksLevel
        {
            int total_bits = 0;
            u64 *d_big = nullptr; // [Q | (Q+1)/2 | Q/q_j ...] words
            Tw *d_invp = nullptr; // (Q/q_j)^-1 mod q_j
        };
        std::map<size_t, CkksLevel> ckks_levels;
        u64 *d_t_mod_q = nullptr;   // [k]: t mod q_i
        u64 t_ratio_lo = 0, t_ratio_hi = 0; // floor(2^128 / t)
        std::vector<std::array<u64, 4>> parms_ids; // parms_ids[L-1] = parms_id of the level with L primes (sb_wire.hpp)
        std::map<uint32_t, uint32_t *> galois_tables; // NTT-form permutation tables (device)
        std::map<size_t, std::shared_ptr<BehzDev>> behz; // per level L
        void *scratch = nullptr;
        int *d_flag = nullptr;               // result word of the range check (op_residues_in_range)
        void *aux_buf = nullptr;             // second grow-only arena (size-3 intermediate of BFV multiply+relinearize)
        size_t aux_bytes = 0;
        size_t scratch_used = 0, aux_used = 0; // bytes of the latest ensure_scratch / ensure_aux request (what wipe_* clears)
        size_t scratch_bytes = 0, table_bytes = 0, scratch_budget = size_t(8) << 30;
        size_t ks_chunk_max = 0;                       // 0 = derived from scratch_budget (sb200_context_set_limit)
        size_t host_stage_bytes = size_t(640) << 20;   // per pipeline slot of the *_host entry points
        LaunchStats stats;
        KsInt ksint;                         // integer key-switching path (sb_ksint.cuh); ready = tables built
        // key switching: 0 = 64-bit digit transforms, 1 = automatic (the integer path from ks_min_digits digits on: below that the
        // L (L+1) large transforms are cheaper than S L + 2 S (L+1) small ones plus the reconstruction), 2 = integer path always
        int ks_algo = 1;
        size_t ks_min_digits = 6;
        bool ksint_on(size_t L) const { return ksint.ready && (ks_algo == 2 || (ks_algo == 1 && L >= ks_min_digits)); }
        IoArena io;
        std::mutex mu;
        // cross-stream ordering of calls that share the scratch arenas (sb_api.cu: StreamOrder)
        cudaEvent_t order_event = nullptr;
        cudaStream_t order_stream = nullptr;
        bool order_valid = false;

        ~Context();
        void *ensure_scratch(size_t bytes);
        void *ensure_aux(size_t bytes);
        // secret-dependent intermediates (decryption phases, encryption noise) do not stay in the arenas: the reference keeps them in
        // clear-on-destruction pools (decryptor.cpp:106-109, util/rlwe.cpp:195, 292)
        void wipe_scratch(cudaStream_t st);
        void wipe_aux(cudaStream_t st);
        const uint32_t *galois_table(uint32_t elt);
        size_t prime_id_aux(size_t aux_index) const { return k + aux_index; }
    };

    std::unique_ptr<Context> make_context(int scheme, size_t n, const u64 *moduli, size_t k, u64 t, int device);

    // drivers (all stream-ordered, no synchronisation); slabs as documented in include/seal_b200.h
    void op_ntt(Context &c, bool inverse, size_t L, size_t size, size_t batch, u64 *d, cudaStream_t st);
    void op_linear(Context &c, int mode, size_t L, size_t size, size_t batch, const u64 *a, const u64 *b, u64 *out, cudaStream_t st);
    void op_multiply_plain(Context &c, size_t L, size_t size, size_t batch, const u64 *a, const u64 *plain, u64 *out, cudaStream_t st);
    // coefficient-form plaintexts [B][n] (words < t): lift + NTT, multiply_plain, add_plain / sub_plain (BFV, BGV);
    // h_cf: per-ciphertext BGV correction factors on the host (nullptr = 1)
    // Decryptor::decrypt (decryptor.cpp:62-197): ct [B][size][L][n] -> CKKS: NTT-form plaintexts [B][L][n]; BFV / BGV:
    // coefficient-form plaintexts [B][n].  h_cf: BGV correction factors per ciphertext (nullptr = 1).
    void secret_key_create(Context &c, const u64 *h_sk, SecretKey &out);
    void op_decrypt(Context &c, SecretKey &sk, size_t L, size_t size, size_t batch, const u64 *ct, const u64 *h_cf, u64 *plain, cudaStream_t st);
    // BatchEncoder::encode / decode (batchencoder.cpp:84-330): values [B][n] (< t) <-> coefficient-form plaintexts [B][n]
    void op_batch_encode(Context &c, size_t batch, const u64 *values, u64 *plain, cudaStream_t st);
    void op_batch_decode(Context &c, size_t batch, const u64 *plain, u64 *values, cudaStream_t st);
    void op_plain_to_ntt(Context &c, size_t L, size_t batch, const u64 *plain, const u64 *h_cf, u64 *out, cudaStream_t st);
    void op_multiply_plain_coeff(Context &c, size_t L, size_t size, size_t batch, bool ct_ntt, const u64 *ct, const u64 *plain, u64 *out,
                                 cudaStream_t st);
    void op_add_plain_coeff(Context &c, size_t L, size_t size, size_t batch, bool subtract, const u64 *ct, const u64 *plain, const u64 *h_cf,
                            u64 *out, cudaStream_t st);
    void op_ckks_multiply(Context &c, size_t L, size_t batch, const u64 *a, const u64 *b, u64 *out3, cudaStream_t st);
    void op_bfv_multiply(Context &c, size_t L, size_t batch, const u64 *a, const u64 *b, u64 *out3, cudaStream_t st);
    // general ciphertext sizes s1 x s2 -> s1+s2-1 (2 x 2 forwards to the specialised kernels above)
    void op_ckks_multiply(Context &c, size_t L, size_t s1, size_t s2, size_t batch, const u64 *a, const u64 *b, u64 *out, cudaStream_t st);
    void op_bfv_multiply(Context &c, size_t L, size_t s1, size_t s2, size_t batch, const u64 *a, const u64 *b, u64 *out, cudaStream_t st);
    void check_sizes(size_t s1, size_t s2);
    void launch_tensor_general(Context &c, const u64 *xa, long long xa_bs, const u64 *xb, long long xb_bs, u64 *out, const int *pid_tab,
                               size_t nb, size_t s1, size_t s2, size_t B, cudaStream_t st);
    void op_relinearize(Context &c, size_t L, size_t batch, const u64 *in3, const KSwitchKey &key, u64 *out2, cudaStream_t st);
    void op_multiply_relinearize(Context &c, size_t L, size_t batch, const u64 *a, const u64 *b, const KSwitchKey &key,
                                 u64 *out2, cudaStream_t st);
    // polys = batch * size polynomials of L components each (any ciphertext size)
    void op_rescale(Context &c, size_t L, size_t polys, const u64 *in, u64 *out, cudaStream_t st);
    void op_mod_switch(Context &c, size_t L, size_t polys, const u64 *in, u64 *out, cudaStream_t st);
    void op_relinearize_sized(Context &c, size_t L, size_t size, size_t batch, const u64 *in, const KSwitchKey &key, u64 *out, cudaStream_t st);
    void op_apply_galois(Context &c, size_t L, size_t batch, const u64 *in2, uint32_t elt, const KSwitchKey &key, u64 *out2,
                         cudaStream_t st);
    size_t keyswitch_chunk(const Context &c, size_t L, size_t batch, bool fused);
    // arithmetic ceilings measured in process: warp-level butterflies (kind 0-2) / multiply-accumulates (kind 3) per second
    double selftest_rate(Context &c, int kind, cudaStream_t st);
    // Ciphertext::expand_seed on the device (sb_prng.cu): seeds [B][8] host words, dst_off [B] host word offsets into d_out of the
    // polynomial ([L][n]) each seed expands into; returns after the expansion has been enqueued and the host arrays are free
    void op_expand_seeded(Context &c, size_t L, size_t B, const u64 *h_seeds, const long long *h_dst_off, u64 *d_out, cudaStream_t st);
    // Encryptor::encrypt_zero_symmetric on the device (sb_prng.cu): bootstrap seeds [B][8] on the host, out [B][2][L][n]
    void op_encrypt_zero_symmetric(Context &c, const SecretKey &sk, size_t L, size_t B, const u64 *h_boot_seeds, bool save_seed, u64 *d_out,
                                   u64 *h_public_seeds, cudaStream_t st);
    // Encryptor(public key)::encrypt_zero on the device (sb_prng.cu): seeds [B][8] on the host, out [B][2][L][n]
    void public_key_create(Context &c, const u64 *h_pk, PublicKey &out);
    void op_encrypt_zero_asymmetric(Context &c, const PublicKey &pk, size_t L, size_t B, const u64 *h_seeds, u64 *d_out, cudaStream_t st);
    // CKKSEncoder::encode / decode (sb_ckks.cu): values = device doubles, [B][count] complex pairs (or reals); plain = [B][L][n] NTT form
    void op_ckks_encode(Context &c, size_t L, size_t B, const double *values, size_t count, bool is_complex, double scale, u64 *plain,
                        cudaStream_t st);
    void op_ckks_decode(Context &c, size_t L, size_t B, const u64 *plain, double scale, double *values, cudaStream_t st);
    const sbh::BehzLevel &behz_host(Context &c, size_t L);
    // wire format support (sb_api.cu): 1 if any residue of data [rows][n] (prime of a row = row % L) is >= its modulus
    bool op_residues_in_range(Context &c, size_t L, size_t rows, const u64 *d, cudaStream_t st);
} // namespace sb
