// seal_b200/csrc/sb_prng.cu -- seed-compressed ciphertexts expanded on the device.
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

    // seeds: [B][8] host words (prng_seed_type); dst_off: [B] host word offsets into d_out of the polynomial to fill ([L][n])
    void op_expand_seeded(Context &c, size_t L, size_t B, const u64 *h_seeds, const long long *h_dst_off, u64 *d_out, cudaStream_t st)
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
            cuda_check(cudaMemcpyAsync(d_seeds, h_seeds, B * 8 * sizeof(u64), cudaMemcpyHostToDevice, st), "seeds H2D");
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
} // namespace sb
