// seal_b200/csrc/sb_prng.cu -- the PRNG-driven pieces on the device: seed-compressed ciphertexts expanded at load time (SURVEY 8f
// rank 3) and symmetric-key encryption of zero (the producer of those ciphertexts; second half of this file).
//
// A ciphertext that comes out of a symmetric-key encryption is saved with its first polynomial only; the second one is
// uniformly random and is carried as the 64-byte seed of the PRNG that produced it (Serializable<Ciphertext>,
// ciphertext.cpp:325-352).  Ciphertext::load re-creates it with Ciphertext::expand_seed (ciphertext.cpp:118-150) ->
// sample_poly_uniform (util/rlwe.cpp:104-132) on a Blake2xbPRNG (randomgen.cpp:204-214).  Doing that here halves the bytes a
// fresh ciphertext moves over PCIe.  What is reproduced, bit for bit:
//   * the PRNG stream: buffer b (4096 bytes) = BLAKE2Xb(out 4096, in = 64-bit counter b, key = seed)  -- RFC 7693 BLAKE2b with
//     the BLAKE2X parameter block: a keyed root hash over the counter, then 64 output blocks B_i = BLAKE2b(root) whose
//     parameter block carries node_offset = i and xof_length = 4096 (util/blake2xb.c);
//   * sample_poly_uniform: the first L*n words of the stream fill the polynomial; every word >= max_multiple of its prime
//     (rejection sampling for uniformity) is replaced, in coefficient order, by the next words of the stream; the survivors are
//     reduced modulo the prime.
// Parallel schedule: one thread per 64-byte output block generates the stream (root hash recomputed per thread: 3 compressions
// per block), one CTA per ciphertext lists the rejected coefficients in order (two passes with a block-wide scan), one thread per
// ciphertext consumes the replacement words (a few thousand sequential steps at n = 65536, 31 primes), one streaming kernel reduces.
#include "sb_engine.cuh"

namespace sb
{
    namespace
    {
        __constant__ u64 c_iv[8] = { 0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                     0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull };
        __constant__ unsigned char c_sigma[12][16] = {
            { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
            { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
            { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
            { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
            { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
            { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
        };
        __device__ __forceinline__ u64 rotr64(u64 v, int s)
        {
            return (v >> s) | (v << (64 - s));
        }
        // BLAKE2b compression function F (RFC 7693 3.2): h updated with one 128-byte block m, byte counter t, last-block flag
        __device__ void blake2b_compress(u64 (&h)[8], const u64 (&m)[16], u64 t, bool last)
        {
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] = h[i], v[i + 8] = c_iv[i];
            v[12] ^= t;
            if (last)
                v[14] = ~v[14];
#define SB_G(a, b, c, d, x, y)                        \
    v[a] += v[b] + (x), v[d] = rotr64(v[d] ^ v[a], 32); \
    v[c] += v[d], v[b] = rotr64(v[b] ^ v[c], 24);       \
    v[a] += v[b] + (y), v[d] = rotr64(v[d] ^ v[a], 16); \
    v[c] += v[d], v[b] = rotr64(v[b] ^ v[c], 63);
            for (int r = 0; r < 12; r++)
            {
                const unsigned char *s = c_sigma[r];
                SB_G(0, 4, 8, 12, m[s[0]], m[s[1]])
                SB_G(1, 5, 9, 13, m[s[2]], m[s[3]])
                SB_G(2, 6, 10, 14, m[s[4]], m[s[5]])
                SB_G(3, 7, 11, 15, m[s[6]], m[s[7]])
                SB_G(0, 5, 10, 15, m[s[8]], m[s[9]])
                SB_G(1, 6, 11, 12, m[s[10]], m[s[11]])
                SB_G(2, 7, 8, 13, m[s[12]], m[s[13]])
                SB_G(3, 4, 9, 14, m[s[14]], m[s[15]])
            }
#undef SB_G
#pragma unroll
            for (int i = 0; i < 8; i++)
                h[i] ^= v[i] ^ v[i + 8];
        }

        constexpr unsigned kXof = 4096;       // UniformRandomGenerator::buffer_size_ (randomgen.h:391)
        constexpr int kWordsPerBuffer = kXof / 8;

        // thread = one 64-byte output block of one PRNG buffer of one ciphertext: W[b][buffer*512 + block*8 .. +8)
        __global__ void __launch_bounds__(128) blake2xb_stream_kernel(const u64 *__restrict__ seeds, u64 *__restrict__ W, long long words_per_ct,
                                                                       long long blocks_per_ct, long long total_blocks)
        {
            const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
            if (g >= total_blocks)
                return;
            const long long b = g / blocks_per_ct, blk = g % blocks_per_ct;
            const u64 counter = static_cast<u64>(blk >> 6); // buffer index = Blake2xbPRNG::counter_ at its refill
            const unsigned node = static_cast<unsigned>(blk & 63);
            u64 h[8], m[16];
            // root: keyed BLAKE2b, digest 64, key length 64, fanout 1, depth 1, xof_length 4096 (blake2xb_init_key)
#pragma unroll
            for (int i = 0; i < 8; i++)
                h[i] = c_iv[i];
            h[0] ^= 64ull | (64ull << 8) | (1ull << 16) | (1ull << 24);
            h[1] ^= static_cast<u64>(kXof) << 32; // bytes 8..11 node_offset = 0, bytes 12..15 xof_length
#pragma unroll
            for (int i = 0; i < 8; i++)
                m[i] = seeds[b * 8 + i], m[i + 8] = 0; // the key, padded to one block
            blake2b_compress(h, m, 128, false);
#pragma unroll
            for (int i = 0; i < 16; i++)
                m[i] = 0;
            m[0] = counter; // the message: the 8-byte counter
            blake2b_compress(h, m, 128 + 8, true);
            // output block `node`: unkeyed BLAKE2b over the 64-byte root with digest 64, fanout 0, depth 0, leaf_length 64,
            // node_offset = node, xof_length 4096, node_depth 0, inner_length 64 (blake2xb_final)
            u64 o[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                o[i] = c_iv[i], m[i] = h[i], m[i + 8] = 0;
            o[0] ^= 64ull | (64ull << 32);
            o[1] ^= static_cast<u64>(node) | (static_cast<u64>(kXof) << 32);
            o[2] ^= 64ull << 8; // byte 16 node_depth = 0, byte 17 inner_length = 64
            blake2b_compress(o, m, 64, true);
            u64 *dst = W + b * words_per_ct + blk * 8;
#pragma unroll
            for (int i = 0; i < 8; i++)
                dst[i] = o[i];
        }

        // one CTA per ciphertext: list the rejected coefficients in order, then let one thread walk the replacement words
        __global__ void __launch_bounds__(1024) seed_reject_kernel(u64 *__restrict__ W, long long words_per_ct, const u64 *__restrict__ max_mult,
                                                                   unsigned *__restrict__ list, long long list_cap, int logn, int L, int *overflow)
        {
            __shared__ unsigned cnt[1024];
            __shared__ unsigned total;
            const int b = blockIdx.x, t = threadIdx.x;
            u64 *w = W + b * words_per_ct;
            unsigned *mine = list + b * list_cap;
            const long long need = static_cast<long long>(L) << logn;
            const long long seg = (need + 1023) / 1024, lo = t * seg, hi = lo + seg < need ? lo + seg : need;
            unsigned c = 0;
            for (long long i = lo; i < hi; i++)
                c += w[i] >= max_mult[i >> logn];
            cnt[t] = c;
            __syncthreads();
            if (t == 0)
            {
                unsigned run = 0;
                for (int i = 0; i < 1024; i++)
                {
                    const unsigned v = cnt[i];
                    cnt[i] = run;
                    run += v;
                }
                total = run;
            }
            __syncthreads();
            unsigned at = cnt[t];
            if (total > list_cap)
            {
                if (t == 0)
                    atomicOr(overflow, 1);
                return;
            }
            for (long long i = lo; i < hi; i++)
                if (w[i] >= max_mult[i >> logn])
                    mine[at++] = static_cast<unsigned>(i);
            __syncthreads();
            if (t != 0)
                return;
            // util/rlwe.cpp:121-128: while (rand >= max_multiple) take the next word of the stream
            long long r = need;
            for (unsigned k = 0; k < total; k++)
            {
                const unsigned idx = mine[k];
                const u64 mm = max_mult[idx >> logn];
                u64 v;
                do
                {
                    if (r >= words_per_ct)
                    {
                        atomicOr(overflow, 1);
                        return;
                    }
                    v = w[r++];
                } while (v >= mm);
                w[idx] = v;
            }
        }

        // barrett_reduce_64 of the accepted words into the second polynomial of each ciphertext
        __global__ void __launch_bounds__(256) seed_reduce_kernel(const u64 *__restrict__ W, long long words_per_ct, u64 *__restrict__ out,
                                                                  const long long *__restrict__ dst_off, const PrimeDev *__restrict__ primes,
                                                                  int logn, int L, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
            if (e >= total)
                return;
            const long long need = static_cast<long long>(L) << logn, b = e / need, i = e % need;
            const PrimeDev P = primes[static_cast<int>(i >> logn)];
            out[dst_off[b] + i] = barrett64(W[b * words_per_ct + i], P.q, P.ratio_hi);
        }
    } // namespace

    // seeds: [B][8] words (prng_seed_type), on the host or (seeds_on_device) already on the device; dst_off: [B] host word offsets into
    // d_out of the polynomial to fill ([L][n])
    static void expand_impl(Context &c, size_t L, size_t B, const u64 *seeds, bool seeds_on_device, const long long *h_dst_off, u64 *d_out,
                            cudaStream_t st)
    {
        if (!B)
            return;
        if (L < 1 || L > c.k)
            throw std::invalid_argument("encrypted is not valid for encryption parameters");
        const long long need = static_cast<long long>(L) * c.n;
        // rejection probability of prime j = (2^64 - max_multiple_j) / 2^64; spare stream words: twice the expectation + a margin
        std::vector<u64> mm(L);
        double expect = 0;
        for (size_t j = 0; j < L; j++)
        {
            const u64 max_random = ~0ull;
            mm[j] = max_random - (max_random % c.q[j]) - 1;
            expect += static_cast<double>(c.n) * (static_cast<double>(max_random - mm[j]) / 18446744073709551616.0);
        }
        for (int attempt = 0; attempt < 4; attempt++)
        {
            const long long spare = static_cast<long long>((2.0 * expect + 2048.0) * (1 << attempt));
            const long long buffers = (need + spare + kWordsPerBuffer - 1) / kWordsPerBuffer, words_per_ct = buffers * kWordsPerBuffer;
            const long long list_cap = spare;
            // scratch: stream words, rejection lists, per-ciphertext seeds / offsets / thresholds, overflow flag
            const size_t bytes = B * words_per_ct * sizeof(u64) + B * list_cap * sizeof(unsigned) + (B * 8 + B + L) * sizeof(u64) + 64;
            u64 *W = static_cast<u64 *>(c.ensure_scratch(bytes));
            unsigned *list = reinterpret_cast<unsigned *>(W + B * words_per_ct);
            u64 *d_seeds = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(list) + ((B * list_cap * sizeof(unsigned) + 15) / 16) * 16);
            long long *d_off = reinterpret_cast<long long *>(d_seeds + B * 8);
            u64 *d_mm = reinterpret_cast<u64 *>(d_off + B);
            int *d_flag = reinterpret_cast<int *>(d_mm + L);
            cuda_check(cudaMemcpyAsync(d_seeds, seeds, B * 8 * sizeof(u64), seeds_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st),
                       "seeds");
            cuda_check(cudaMemcpyAsync(d_off, h_dst_off, B * sizeof(long long), cudaMemcpyHostToDevice, st), "offsets H2D");
            cuda_check(cudaMemcpyAsync(d_mm, mm.data(), L * sizeof(u64), cudaMemcpyHostToDevice, st), "thresholds H2D");
            cuda_check(cudaMemsetAsync(d_flag, 0, sizeof(int), st), "memset");
            const long long blocks_per_ct = words_per_ct / 8, total_blocks = blocks_per_ct * static_cast<long long>(B);
            c.stats.begin("seed_stream", 0, 8.0 * B * words_per_ct, st);
            blake2xb_stream_kernel<<<static_cast<unsigned>((total_blocks + 127) / 128), 128, 0, st>>>(d_seeds, W, words_per_ct, blocks_per_ct, total_blocks);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "blake2xb_stream_kernel");
            c.stats.begin("seed_reject", 0, 16.0 * B * need, st);
            seed_reject_kernel<<<static_cast<unsigned>(B), 1024, 0, st>>>(W, words_per_ct, d_mm, list, list_cap, c.logn, static_cast<int>(L), d_flag);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "seed_reject_kernel");
            int flag = 0;
            cuda_check(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "flag D2H");
            cuda_check(cudaStreamSynchronize(st), "synchronize"); // also: the host arrays above may now go out of scope
            if (flag)
                continue; // more rejections than spare words (probability far below 2^-40 at the first attempt): take more
            const long long total = need * static_cast<long long>(B);
            c.stats.begin("seed_reduce", 0, 16.0 * total, st);
            seed_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(W, words_per_ct, d_out, d_off, c.d_primes, c.logn,
                                                                                           static_cast<int>(L), total);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "seed_reduce_kernel");
            return;
        }
        throw std::logic_error("seed expansion ran out of random words");
    }

    void op_expand_seeded(Context &c, size_t L, size_t B, const u64 *h_seeds, const long long *h_dst_off, u64 *d_out, cudaStream_t st)
    {
        expand_impl(c, L, B, h_seeds, false, h_dst_off, d_out, st);
    }

    // ------------------------------------------------------------------------------------------------------------------------------
    // Encryptor::encrypt_zero_symmetric (encryptor.cpp:168-173 -> util/rlwe.cpp:264-408) for a batch of ciphertexts.
    // Per ciphertext, from ONE bootstrap PRNG (Blake2xb of a fresh 64-byte seed):
    //   bytes [0, 64)          the public seed P: c_1 = sample_poly_uniform(Blake2xb(P))            (rlwe.cpp:298-324)
    //   bytes [64 + 6i, +6)    the noise coefficient i: hw(x0, x1, x2 & 31) - hw(x3, x4, x5 & 31)    (sample_poly_cbd, rlwe.cpp:73-102)
    //   c_0 = -(c_1 s + e)  [CKKS: everything in NTT form;  BGV: -(c_1 s + t e), NTT form;  BFV: coefficient form, see below]
    // BFV ciphertexts are in coefficient form (rlwe.cpp:314-324, 347-372): when the seed is kept (Serializable<Ciphertext>) the sampled
    // polynomial IS c_1 and a transformed copy is multiplied with the key; otherwise the sample is taken as NTT(c_1).
    // The uniform polynomial reuses the expansion above with the seeds left on the device; the noise costs 6 bytes per coefficient and
    // is re-read per prime from L2; the transforms are the library's own (op_ntt).
    namespace
    {
        __device__ __forceinline__ int cbd_noise(const unsigned char *boot, long long i)
        {
            const unsigned short *x = reinterpret_cast<const unsigned short *>(boot + 64 + 6 * i); // 64 + 6 i is even
            const unsigned w0 = x[0], w1 = x[1], w2 = x[2];
            const unsigned lo = (w0 | (w1 << 16)) & 0x1FFFFFu, hi = ((w1 >> 8) | (w2 << 8)) & 0x1FFFFFu;
            return __popc(lo) - __popc(hi);
        }
        __device__ __forceinline__ u64 neg_sum(u64 a, u64 b, u64 q) // -(a + b) mod q, operands < q
        {
            u64 v = a + b;
            v -= v >= q ? q : 0;
            return v ? q - v : 0;
        }
        // MODE 0  tmp = m_j e                 (CKKS m = 1 / BGV m = t mod q_j; transformed afterwards)
        // MODE 1  c_0 = -(s c_1 + tmp)        (NTT form schemes)
        // MODE 2  tmp = c_1                   (BFV with the seed kept: the copy that gets transformed)
        // MODE 3  tmp = s tmp
        // MODE 4  c_0 = -(tmp + e)
        // MODE 5  c_0 = s c_1                 (BFV without the seed: both polynomials are inverse-transformed afterwards)
        // MODE 6  c_0 = -(c_0 + e)
        template <int MODE>
        __global__ void __launch_bounds__(256) enc_kernel(u64 *__restrict__ out, u64 *__restrict__ tmp, const u64 *__restrict__ sk,
                                                          const unsigned char *__restrict__ boot, long long boot_bytes, const u64 *__restrict__ mult,
                                                          const PrimeDev *__restrict__ primes, int logn, int L, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * L * n
            if (e >= total)
                return;
            const long long poly = static_cast<long long>(L) << logn, b = e / poly, r = e % poly;
            const int j = static_cast<int>(r >> logn);
            const long long i = r & ((1ll << logn) - 1);
            const PrimeDev P = primes[j];
            u64 *c0 = out + b * 2 * poly + r, *c1 = c0 + poly;
            u64 noise = 0;
            if (MODE == 0 || MODE == 4 || MODE == 6)
            {
                const int v = cbd_noise(boot + b * boot_bytes, i);
                noise = v >= 0 ? static_cast<u64>(v) : P.q - static_cast<u64>(-v);
            }
            if (MODE == 0)
                tmp[e] = mult ? mulmod_barrett(noise, mult[j], P) : noise;
            else if (MODE == 1)
                *c0 = neg_sum(mulmod_barrett(sk[r], *c1, P), tmp[e], P.q);
            else if (MODE == 2)
                tmp[e] = *c1;
            else if (MODE == 3)
                tmp[e] = mulmod_barrett(sk[r], tmp[e], P);
            else if (MODE == 4)
                *c0 = neg_sum(tmp[e], noise, P.q);
            else if (MODE == 5)
                *c0 = mulmod_barrett(sk[r], *c1, P);
            else
                *c0 = neg_sum(*c0, noise, P.q);
        }
    } // namespace

    // h_boot_seeds [B][8]: the seeds of the bootstrap PRNGs (one fresh seed per ciphertext); d_out [B][2][L][n];
    // h_public_seeds: nullptr or [B][8], receives the seed c_1 of each ciphertext expands from (what Serializable<Ciphertext> stores)
    void op_encrypt_zero_symmetric(Context &c, const SecretKey &sk, size_t L, size_t B, const u64 *h_boot_seeds, bool save_seed, u64 *d_out,
                                   u64 *h_public_seeds, cudaStream_t st)
    {
        if (!B)
            return;
        if (L < 1 || L > c.k || (c.k > 1 && L == c.k))
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (!sk.d_pow || sk.ctx != &c)
            throw std::invalid_argument("secret key is not valid for encryption parameters");
        const long long poly = static_cast<long long>(L) * c.n, total = poly * static_cast<long long>(B);
        // bootstrap stream: whole PRNG buffers covering 64 + 6 n bytes
        const long long boot_words = ((64 + 6 * static_cast<long long>(c.n) + kXof - 1) / kXof) * kWordsPerBuffer, boot_bytes = boot_words * 8;
        const size_t bytes = B * boot_bytes + static_cast<size_t>(total) * sizeof(u64) + (B * 8 + B * 8 + L) * sizeof(u64);
        u64 *Wb = static_cast<u64 *>(c.ensure_aux(bytes));
        u64 *tmp = Wb + B * boot_words, *d_boot_seeds = tmp + total, *d_pub = d_boot_seeds + B * 8, *d_mult = d_pub + B * 8;
        cuda_check(cudaMemcpyAsync(d_boot_seeds, h_boot_seeds, B * 8 * sizeof(u64), cudaMemcpyHostToDevice, st), "bootstrap seeds H2D");
        const u64 *mult = nullptr;
        std::vector<u64> h_mult;
        if (c.scheme == 3 /* BGV */)
        {
            h_mult.resize(L);
            for (size_t j = 0; j < L; j++)
                h_mult[j] = c.t % c.q[j];
            cuda_check(cudaMemcpyAsync(d_mult, h_mult.data(), L * sizeof(u64), cudaMemcpyHostToDevice, st), "t mod q H2D");
            mult = d_mult;
        }
        const long long blocks_per_ct = boot_words / 8, total_blocks = blocks_per_ct * static_cast<long long>(B);
        c.stats.begin("enc_bootstrap_stream", 0, 8.0 * B * boot_words, st);
        blake2xb_stream_kernel<<<static_cast<unsigned>((total_blocks + 127) / 128), 128, 0, st>>>(d_boot_seeds, Wb, boot_words, blocks_per_ct, total_blocks);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "blake2xb_stream_kernel");
        cuda_check(cudaMemcpy2DAsync(d_pub, 64, Wb, boot_bytes, 64, B, cudaMemcpyDeviceToDevice, st), "public seeds");
        std::vector<long long> off(B);
        for (size_t b = 0; b < B; b++)
            off[b] = static_cast<long long>(b) * 2 * poly + poly;
        expand_impl(c, L, B, d_pub, true, off.data(), d_out, st); // c_1 (synchronises the stream: the host vectors above are free)
        if (h_public_seeds)
            cuda_check(cudaMemcpyAsync(h_public_seeds, d_pub, B * 8 * sizeof(u64), cudaMemcpyDeviceToHost, st), "public seeds D2H");
        const unsigned grid = static_cast<unsigned>((total + 255) / 256);
        const unsigned char *boot = reinterpret_cast<const unsigned char *>(Wb);
        const u64 *s = sk.d_pow; // s^1, [k][n] at the key level: the first L rows are the level's primes
        const int logn = c.logn, Li = static_cast<int>(L);
#define SB_ENC(MODE, name)                                                                                        \
    c.stats.begin(name, 0, 24.0 * total, st);                                                                      \
    enc_kernel<MODE><<<grid, 256, 0, st>>>(d_out, tmp, s, boot, boot_bytes, mult, c.d_primes, logn, Li, total);    \
    c.stats.end(st);                                                                                               \
    cuda_check(cudaGetLastError(), name);
        if (c.scheme != 1 /* CKKS, BGV */)
        {
            SB_ENC(0, "enc_noise")
            op_ntt(c, false, L, 1, B, tmp, st);
            SB_ENC(1, "enc_combine")
        }
        else if (save_seed)
        {
            SB_ENC(2, "enc_copy")
            op_ntt(c, false, L, 1, B, tmp, st);
            SB_ENC(3, "enc_mul_key")
            op_ntt(c, true, L, 1, B, tmp, st);
            SB_ENC(4, "enc_combine")
        }
        else
        {
            SB_ENC(5, "enc_mul_key")
            op_ntt(c, true, L, 2, B, d_out, st);
            SB_ENC(6, "enc_combine")
        }
#undef SB_ENC
        c.wipe_aux(st); // bootstrap stream and noise: with c_0 they give c_1 s (util/rlwe.cpp:292: clear-on-destruction pool)
        if (h_public_seeds)
            cuda_check(cudaStreamSynchronize(st), "synchronize");
    }

    // ------------------------------------------------------------------------------------------------------------------------------
    // Encryptor::encrypt_zero with a public key (encryptor.cpp:88-174 -> util::encrypt_zero_asymmetric, util/rlwe.cpp:184-276).
    // One PRNG per ciphertext yields, in order: the ternary polynomial u -- one 32-bit word r per coefficient mapped to (r * 3) >> 32
    // (std::uniform_int_distribution<uint64_t>(0, 2) over a 32-bit generator in libstdc++ >= 11: Lemire's method, redraw only for
    // r == 0) -- then the noise polynomials e_0, e_1 (6 bytes per coefficient each).  c_j = pk_j u + e_j (BGV: t e_j) is formed one
    // level above the requested one (L + 1 primes, the level's prev_context_data) and divided down by that level's last prime with
    // the evaluator's own kernels (CKKS: op_rescale, BFV / BGV: op_mod_switch); at the key level (L == k) it is returned as sampled.
    // A redraw shifts everything behind it by one word: the parallel kernel only counts zero words, and a one-thread-per-ciphertext
    // kernel re-walks the (probability 2^-32 per coefficient) ciphertexts that had one.
    namespace
    {
        __device__ __forceinline__ void ternary_store(u64 *__restrict__ u, long long b, long long i, unsigned r, const PrimeDev *__restrict__ primes,
                                                      int logn, int Lp)
        {
            const u64 v = (static_cast<u64>(r) * 3u) >> 32;
            u64 *dst = u + ((b * Lp) << logn) + i;
            for (int j = 0; j < Lp; j++)
                dst[static_cast<long long>(j) << logn] = v + (v == 0 ? primes[j].q : 0) - 1; // rlwe.cpp:32-37
        }
        __global__ void __launch_bounds__(256) enc_ternary_kernel(const u64 *__restrict__ W, long long words_per_ct, u64 *__restrict__ u,
                                                                  unsigned *__restrict__ pos, const PrimeDev *__restrict__ primes, int logn, int Lp,
                                                                  long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * n
            if (e >= total)
                return;
            const long long b = e >> logn, i = e & ((1ll << logn) - 1);
            const unsigned r = reinterpret_cast<const unsigned *>(W + b * words_per_ct)[i];
            if (r == 0)
                atomicAdd(pos + b, 1u); // pos[b] starts at n: anything else marks the ciphertext for the sequential walk
            ternary_store(u, b, i, r, primes, logn, Lp);
        }
        __global__ void enc_ternary_fix_kernel(const u64 *__restrict__ W, long long words_per_ct, u64 *__restrict__ u, unsigned *__restrict__ pos,
                                               const PrimeDev *__restrict__ primes, int logn, int Lp, int *overflow)
        {
            const long long b = blockIdx.x, n = 1ll << logn;
            if (threadIdx.x != 0 || pos[b] == static_cast<unsigned>(n))
                return;
            const unsigned *w = reinterpret_cast<const unsigned *>(W + b * words_per_ct);
            const long long cap = words_per_ct * 2;
            long long p = 0;
            for (long long i = 0; i < n; i++)
            {
                unsigned r;
                do
                {
                    if (p >= cap)
                    {
                        atomicOr(overflow, 1);
                        return;
                    }
                    r = w[p++];
                } while (r == 0);
                ternary_store(u, b, i, r, primes, logn, Lp);
            }
            pos[b] = static_cast<unsigned>(p);
        }
        // MODE 0  T_p = m_j e_p                  (CKKS / BGV: transformed afterwards)
        // MODE 1  T_p = u pk_p + T_p             (CKKS / BGV)
        // MODE 2  T_p = u pk_p                   (BFV: inverse-transformed afterwards)
        // MODE 3  T_p = T_p + e_p                (BFV)
        template <int MODE>
        __global__ void __launch_bounds__(256) enc_pk_kernel(u64 *__restrict__ T, const u64 *__restrict__ u, const u64 *__restrict__ pk, long long pk_poly,
                                                             const u64 *__restrict__ W, long long words_per_ct, const unsigned *__restrict__ pos,
                                                             const u64 *__restrict__ mult, const PrimeDev *__restrict__ primes, int logn, int Lp,
                                                             long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * 2 * Lp * n
            if (e >= total)
                return;
            const long long n = 1ll << logn, poly = static_cast<long long>(Lp) << logn;
            const long long b = e / (2 * poly), rem = e % (2 * poly), p = rem / poly, r = rem % poly;
            const int j = static_cast<int>(r >> logn);
            const long long i = r & (n - 1);
            const PrimeDev P = primes[j];
            u64 noise = 0;
            if (MODE == 0 || MODE == 3)
            {
                // the noise bytes start behind the ternary words: byte 4 pos[b]; polynomial p, coefficient i at + 6 (p n + i);
                // cbd_noise reads from its argument + 64 (the symmetric layout), hence the - 64
                const unsigned char *base = reinterpret_cast<const unsigned char *>(W + b * words_per_ct) + 4ll * pos[b] - 64;
                const int v = cbd_noise(base, p * n + i);
                noise = v >= 0 ? static_cast<u64>(v) : P.q - static_cast<u64>(-v);
            }
            if (MODE == 0)
                T[e] = mult ? mulmod_barrett(noise, mult[j], P) : noise;
            else
            {
                u64 v = 0;
                if (MODE == 1 || MODE == 2)
                    v = mulmod_barrett(u[b * poly + r], pk[p * pk_poly + r], P);
                if (MODE == 1)
                    v += T[e];
                if (MODE == 3)
                    v = T[e] + noise;
                T[e] = v >= P.q ? v - P.q : v;
            }
        }
    } // namespace

    void public_key_create(Context &c, const u64 *h_pk, PublicKey &out)
    {
        const size_t words = 2 * c.k * c.n;
        std::vector<u64> rows(c.k);
        cuda_check(cudaMalloc(&out.d_key, words * sizeof(u64)), "cudaMalloc(public key)");
        out.ctx = &c;
        cuda_check(cudaMemcpy(out.d_key, h_pk, words * sizeof(u64), cudaMemcpyHostToDevice), "upload public key");
        if (!op_residues_in_range(c, c.k, 2 * c.k, out.d_key, nullptr))
        {
            cudaFree(out.d_key);
            out.d_key = nullptr;
            throw std::invalid_argument("public key is not valid for encryption parameters");
        }
    }

    // h_seeds [B][8]: the seed of each ciphertext's PRNG; d_out [B][2][L][n] at the level with L primes (L == k: the key level)
    void op_encrypt_zero_asymmetric(Context &c, const PublicKey &pk, size_t L, size_t B, const u64 *h_seeds, u64 *d_out, cudaStream_t st)
    {
        if (!B)
            return;
        if (L < 1 || L > c.k)
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (!pk.d_key || pk.ctx != &c)
            throw std::invalid_argument("public key is not valid for encryption parameters");
        const size_t Lp = L < c.k ? L + 1 : L;
        const long long n = static_cast<long long>(c.n), poly = static_cast<long long>(Lp) * n, total1 = static_cast<long long>(B) * poly;
        // stream: 4 n bytes of ternary words + 12 n bytes of noise + spare words for redraws, in whole PRNG buffers
        const long long words = ((16 * n + 256 + kXof - 1) / kXof) * kWordsPerBuffer;
        const size_t bytes = (B * words + static_cast<size_t>(total1) * 3 + B * 8 + Lp) * sizeof(u64) + (B + 4) * sizeof(unsigned);
        u64 *W = static_cast<u64 *>(c.ensure_aux(bytes));
        u64 *U = W + B * words, *T = U + total1, *d_seeds = T + 2 * total1, *d_mult = d_seeds + B * 8;
        unsigned *d_pos = reinterpret_cast<unsigned *>(d_mult + Lp);
        int *d_over = reinterpret_cast<int *>(d_pos + B);
        u64 *dst = Lp == L ? d_out : T;
        cuda_check(cudaMemcpyAsync(d_seeds, h_seeds, B * 8 * sizeof(u64), cudaMemcpyHostToDevice, st), "seeds H2D");
        std::vector<unsigned> h_pos(B, static_cast<unsigned>(n));
        cuda_check(cudaMemcpyAsync(d_pos, h_pos.data(), B * sizeof(unsigned), cudaMemcpyHostToDevice, st), "positions H2D");
        cuda_check(cudaMemsetAsync(d_over, 0, sizeof(int), st), "memset");
        const u64 *mult = nullptr;
        std::vector<u64> h_mult;
        if (c.scheme == 3 /* BGV */)
        {
            h_mult.resize(Lp);
            for (size_t j = 0; j < Lp; j++)
                h_mult[j] = c.t % c.q[j];
            cuda_check(cudaMemcpyAsync(d_mult, h_mult.data(), Lp * sizeof(u64), cudaMemcpyHostToDevice, st), "t mod q H2D");
            mult = d_mult;
        }
        const long long blocks_per_ct = words / 8, total_blocks = blocks_per_ct * static_cast<long long>(B);
        c.stats.begin("enc_stream", 0, 8.0 * B * words, st);
        blake2xb_stream_kernel<<<static_cast<unsigned>((total_blocks + 127) / 128), 128, 0, st>>>(d_seeds, W, words, blocks_per_ct, total_blocks);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "blake2xb_stream_kernel");
        const long long bn = static_cast<long long>(B) * n;
        const int logn = c.logn, Lpi = static_cast<int>(Lp);
        c.stats.begin("enc_ternary", 0, (4.0 + 8.0 * Lp) * bn, st);
        enc_ternary_kernel<<<static_cast<unsigned>((bn + 255) / 256), 256, 0, st>>>(W, words, U, d_pos, c.d_primes, logn, Lpi, bn);
        enc_ternary_fix_kernel<<<static_cast<unsigned>(B), 32, 0, st>>>(W, words, U, d_pos, c.d_primes, logn, Lpi, d_over);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "enc_ternary_kernel");
        op_ntt(c, false, Lp, 1, B, U, st);
        const long long total2 = 2 * total1, pk_poly = static_cast<long long>(c.k) * n;
        const unsigned grid = static_cast<unsigned>((total2 + 255) / 256);
#define SB_ENC(MODE, name)                                                                                                          \
    c.stats.begin(name, 0, 24.0 * total2, st);                                                                                       \
    enc_pk_kernel<MODE><<<grid, 256, 0, st>>>(dst, U, pk.d_key, pk_poly, W, words, d_pos, mult, c.d_primes, logn, Lpi, total2);      \
    c.stats.end(st);                                                                                                                 \
    cuda_check(cudaGetLastError(), name);
        if (c.scheme != 1 /* CKKS, BGV */)
        {
            SB_ENC(0, "enc_noise")
            op_ntt(c, false, Lp, 2, B, dst, st);
            SB_ENC(1, "enc_combine")
        }
        else
        {
            SB_ENC(2, "enc_mul_key")
            op_ntt(c, true, Lp, 2, B, dst, st);
            SB_ENC(3, "enc_combine")
        }
#undef SB_ENC
        int over = 0;
        cuda_check(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st), "flag D2H");
        cuda_check(cudaStreamSynchronize(st), "synchronize"); // also: the host vectors above may now go out of scope
        if (over)
            throw std::logic_error("ternary sampling ran out of random words");
        if (Lp != L)
        {
            // Encryptor::encrypt_zero_internal (encryptor.cpp:125-161): switch down to the requested level
            if (c.scheme == 2)
                op_rescale(c, Lp, 2 * B, T, d_out, st);
            else
                op_mod_switch(c, Lp, 2 * B, T, d_out, st);
        }
        c.wipe_aux(st); // u, the noise and the PRNG stream (util/rlwe.cpp:195: clear-on-destruction pool)
    }
} // namespace sb
