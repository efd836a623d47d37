// seal_b200/csrc/sb_ckks.cu -- CKKSEncoder::encode / decode on the device (ckks.h:455-807), the producer / consumer of the
// NTT-form plaintexts the evaluator works on.
//
// Floating point with an integer contract: the reference runs a double-precision complex FFT (util/dwthandler.h) over the slot
// values, rounds every coefficient to an integer and reduces it modulo each prime, so a plaintext is only reproducible bit for
// bit if every floating-point operation is the reference's, in the reference's order.  Each butterfly below is therefore written
// with the round-to-nearest intrinsics (__dadd_rn / __dsub_rn / __dmul_rn: never contracted into FMA), a complex product is
// (ac - bd, ad + bc) exactly as std::complex multiplies, the scaling by scale / n is merged into the last stage as
// DWTHandler::transform_from_rev does, and the root tables come from the host (sbh::ckks_tables: std::polar on one eighth of the
// circle + symmetry, util/croots.cpp).  The butterflies of one stage are independent, so the parallel schedule does not change
// any value.  After the transform the three decomposition branches of the reference (coefficients below 2^64, below 2^128,
// multi-precision) all reduce the same exact integer |round(c)| = mantissa * 2^exponent; one kernel does that for every size.
//
// Schedule: the stages with gap < 2^kFftLocalLog run in shared memory on contiguous blocks of 2^(kFftLocalLog+1)... (one CTA per block
// of 2048 points, 1024 threads, one butterfly per thread per stage); the remaining log n - 11 stages are one streaming kernel
// each.  At n = 65536 that is 6 passes over 1 MB per plaintext: noise next to the 31 NTTs that follow.
#include "sb_engine.cuh"
#include <cmath>
#include <cstring>

namespace sb
{
    namespace
    {
        constexpr int kFftLocalLog = 11; // points per CTA of the shared-memory part: 2^11 complex doubles = 32 KB

        __device__ __forceinline__ double2 cadd(double2 a, double2 b)
        {
            return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
        }
        __device__ __forceinline__ double2 csubd(double2 a, double2 b)
        {
            return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));
        }
        __device__ __forceinline__ double2 cmul(double2 a, double2 r) // std::complex product: (ac - bd, ad + bc)
        {
            return make_double2(__dsub_rn(__dmul_rn(a.x, r.x), __dmul_rn(a.y, r.y)), __dadd_rn(__dmul_rn(a.x, r.y), __dmul_rn(a.y, r.x)));
        }
        __device__ __forceinline__ double2 cscale(double2 a, double s)
        {
            return make_double2(__dmul_rn(a.x, s), __dmul_rn(a.y, s));
        }

        // values [B][count] (complex: (re, im) pairs; real: one double) -> V[b][map[i]] = v, V[b][map[i + slots]] = conj(v)
        __global__ void __launch_bounds__(256) ckks_scatter_kernel(const double *__restrict__ values, int is_complex, long long count,
                                                                   const uint32_t *__restrict__ map, double2 *__restrict__ V, int logn,
                                                                   long long total, int *flags)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
            if (e >= total)
                return;
            const long long b = e / count, i = e % count;
            const long long n = 1ll << logn, slots = n >> 1;
            double re, im = 0.0;
            if (is_complex)
                re = values[2 * e], im = values[2 * e + 1];
            else
                re = values[e];
            if (!isfinite(re) || !isfinite(im))
                atomicOr(flags, 1); // "values must be finite"
            V[b * n + map[i]] = make_double2(re, im);
            V[b * n + map[i + slots]] = make_double2(re, -im);
        }

        // DWTHandler::transform_from_rev, the stages with gap < 2^local (gap = 1, 2, ...): stage with gap g has m = n / 2g groups, group i
        // uses inv_roots[n - 2m + i + 1]; the stage with m == 1 (only reached here when n <= 2^kFftLocalLog) carries the scalar
        __global__ void __launch_bounds__(1024) ckks_ifft_local_kernel(double2 *__restrict__ V, const double2 *__restrict__ inv_roots, int logn,
                                                                       int local, double scalar)
        {
            extern __shared__ double2 sm[];
            const long long n = 1ll << logn;
            const int pts = 1 << local, t = threadIdx.x;
            const long long blocks_per_poly = n >> local;
            const long long poly = blockIdx.x / blocks_per_poly, blk = blockIdx.x % blocks_per_poly;
            double2 *base = V + poly * n + blk * pts;
            for (int i = t; i < pts; i += blockDim.x)
                sm[i] = base[i];
            __syncthreads();
            for (int s = 0; s < local; s++)
            {
                const int gap = 1 << s;
                const long long m = n >> (s + 1);
                for (int w = t; w < pts / 2; w += blockDim.x)
                {
                    const int g = w >> s, j = w & (gap - 1), xi = (g << (s + 1)) + j;
                    const long long gi = (blk << (local - s - 1)) + g; // group index within the polynomial
                    const double2 r = inv_roots[n - 2 * m + gi + 1];
                    const double2 u = sm[xi], v = sm[xi + gap];
                    if (m > 1)
                    {
                        sm[xi] = cadd(u, v);
                        sm[xi + gap] = cmul(csubd(u, v), r);
                    }
                    else
                    {
                        sm[xi] = cscale(cadd(u, v), scalar);
                        sm[xi + gap] = cmul(csubd(u, v), cscale(r, scalar));
                    }
                }
                __syncthreads();
            }
            for (int i = t; i < pts; i += blockDim.x)
                base[i] = sm[i];
        }

        // one stage (gap = 2^s) over global memory; LAST: the scaled stage (m == 1)
        template <bool LAST>
        __global__ void __launch_bounds__(256) ckks_ifft_stage_kernel(double2 *__restrict__ V, const double2 *__restrict__ inv_roots, int logn, int s,
                                                                      double scalar, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * n / 2 butterflies
            if (e >= total)
                return;
            const long long n = 1ll << logn, half = n >> 1, poly = e / half, w = e % half;
            const long long gap = 1ll << s, m = n >> (s + 1), g = w >> s, j = w & (gap - 1);
            double2 *x = V + poly * n + (g << (s + 1)) + j, *y = x + gap;
            const double2 r = inv_roots[n - 2 * m + g + 1];
            const double2 u = *x, v = *y;
            if (!LAST)
            {
                *x = cadd(u, v);
                *y = cmul(csubd(u, v), r);
            }
            else
            {
                *x = cscale(cadd(u, v), scalar);
                *y = cmul(csubd(u, v), cscale(r, scalar));
            }
        }

        // max |Re c| (as the bit pattern of a non-negative double: ordered like the integers) and the NaN flag (ckks.h:525-545)
        __global__ void __launch_bounds__(256) ckks_max_kernel(const double2 *__restrict__ V, long long total, unsigned long long *max_bits, int *flags)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
            double a = 0.0;
            if (e < total)
                a = fabs(V[e].x);
            if (isnan(a))
            {
                atomicOr(flags, 2);
                a = 0.0;
            }
            unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(a));
            for (int o = 16; o; o >>= 1)
            {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
                bits = other > bits ? other : bits;
            }
            if ((threadIdx.x & 31) == 0 && bits)
                atomicMax(max_bits, bits);
        }

        // round(Re c) -> residues modulo every prime of the level (ckks.h:556-668): |round(c)| = mant * 2^ex exactly
        __global__ void __launch_bounds__(256) ckks_reduce_kernel(const double2 *__restrict__ V, u64 *__restrict__ out, const PrimeDev *__restrict__ primes,
                                                                  int logn, int L, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * n
            if (e >= total)
                return;
            const long long n = 1ll << logn, b = e >> logn, i = e & (n - 1);
            double d = round(V[e].x); // half away from zero, like std::round
            const bool negative = signbit(d);
            d = fabs(d);
            u64 mant = 0;
            int ex = 0;
            if (d >= 1.0)
            {
                if (d < 18446744073709551616.0)
                    mant = __double2ull_rz(d);
                else
                {
                    const long long bits = __double_as_longlong(d);
                    ex = static_cast<int>((bits >> 52) & 0x7ff) - 1075; // d = (2^52 + fraction) * 2^ex, ex >= 12 here
                    mant = (static_cast<u64>(bits) & 0xFFFFFFFFFFFFFull) | (1ull << 52);
                }
            }
            u64 *dst = out + b * L * n + i;
            for (int j = 0; j < L; j++)
            {
                const PrimeDev P = primes[j];
                u64 r = barrett64(mant, P.q, P.ratio_hi);
                for (int left = ex; left > 0; left -= 32) // times 2^ex, 32 bits at a time (large coefficients only)
                    r = mulmod_barrett(r, 1ull << (left >= 32 ? 32 : left), P);
                dst[static_cast<long long>(j) * n] = negative && r ? P.q - r : r;
            }
        }

        // ---- decode ----
        // RNSBase::compose (rns.cpp:321-352) of one coefficient + the conversion to double (ckks.h:740-778).
        // big = [Q (L) | threshold (L) | punctured (L * L)] words; one thread per coefficient, the L-word integers in local memory
        constexpr int kMaxWords = 64; // SEAL_COEFF_MOD_COUNT_MAX (util/defines.h)
        __global__ void __launch_bounds__(128) ckks_compose_kernel(const u64 *__restrict__ coef, const u64 *__restrict__ big, const Tw *__restrict__ invp,
                                                                   const PrimeDev *__restrict__ primes, double2 *__restrict__ V, int logn, int L,
                                                                   double inv_scale, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * n
            if (e >= total)
                return;
            const long long n = 1ll << logn, b = e >> logn, i = e & (n - 1);
            const u64 *Q = big, *thr = big + L, *punct = big + 2 * L;
            u64 val[kMaxWords];
            for (int w = 0; w < L; w++)
                val[w] = 0;
            const u64 *src = coef + b * L * n + i;
            if (L == 1)
                val[0] = src[0];
            else
                for (int j = 0; j < L; j++)
                {
                    const u64 t = mul_shoup(src[static_cast<long long>(j) * n], invp[j], primes[j].q);
                    // val += punct_j * t  (punct_j * t < Q), then one conditional subtraction of Q
                    u64 carry = 0, cy = 0;
                    for (int w = 0; w < L; w++)
                    {
                        const u64 p = punct[j * L + w];
                        const u64 lo = p * t, hi = __umul64hi(p, t);
                        u64 term = lo + carry;
                        carry = hi + (term < lo);
                        u64 sum = val[w] + term;
                        const u64 c1 = sum < term;
                        sum += cy;
                        cy = c1 | (sum < cy);
                        val[w] = sum;
                    }
                    bool ge = cy != 0;
                    if (!ge)
                    {
                        ge = true;
                        for (int w = L - 1; w >= 0; w--)
                            if (val[w] != Q[w])
                            {
                                ge = val[w] > Q[w];
                                break;
                            }
                    }
                    if (ge)
                    {
                        u64 bw = 0;
                        for (int w = 0; w < L; w++)
                        {
                            const u64 a = val[w], s1 = a - Q[w], s2 = s1 - bw;
                            bw = (a < Q[w]) | (s1 < bw);
                            val[w] = s2;
                        }
                    }
                }
            bool upper = true; // is_greater_than_or_equal_uint(value, upper_half_threshold)
            for (int w = L - 1; w >= 0; w--)
                if (val[w] != thr[w])
                {
                    upper = val[w] > thr[w];
                    break;
                }
            // the reference accumulates word by word in doubles, the words of Q - value taken WORDWISE (no borrow): keep that
            double acc = 0.0, s64 = inv_scale;
            for (int w = 0; w < L; w++, s64 = __dmul_rn(s64, 18446744073709551616.0))
            {
                if (upper)
                {
                    if (val[w] > Q[w])
                    {
                        const u64 diff = val[w] - Q[w];
                        acc = __dadd_rn(acc, __dmul_rn(__ull2double_rn(diff), s64));
                    }
                    else
                    {
                        const u64 diff = Q[w] - val[w];
                        acc = __dsub_rn(acc, diff ? __dmul_rn(__ull2double_rn(diff), s64) : 0.0);
                    }
                }
                else
                    acc = __dadd_rn(acc, val[w] ? __dmul_rn(__ull2double_rn(val[w]), s64) : 0.0);
            }
            V[e] = make_double2(acc, 0.0);
        }

        // DWTHandler::transform_to_rev: stage with m groups (gap = n / 2m), group i uses roots[m + i]; u, v = y r; x = u + v, y = u - v
        __global__ void __launch_bounds__(256) ckks_fft_stage_kernel(double2 *__restrict__ V, const double2 *__restrict__ roots, int logn, int s,
                                                                     long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
            if (e >= total)
                return;
            const long long n = 1ll << logn, half = n >> 1, poly = e / half, w = e % half;
            const long long gap = 1ll << s, m = n >> (s + 1), g = w >> s, j = w & (gap - 1);
            double2 *x = V + poly * n + (g << (s + 1)) + j, *y = x + gap;
            const double2 u = *x, v = cmul(*y, roots[m + g]);
            *x = cadd(u, v);
            *y = csubd(u, v);
        }
        // the stages with gap < 2^local (gap = 2^(local-1) ... 1) in shared memory
        __global__ void __launch_bounds__(1024) ckks_fft_local_kernel(double2 *__restrict__ V, const double2 *__restrict__ roots, int logn, int local)
        {
            extern __shared__ double2 sm[];
            const long long n = 1ll << logn;
            const int pts = 1 << local, t = threadIdx.x;
            const long long blocks_per_poly = n >> local;
            const long long poly = blockIdx.x / blocks_per_poly, blk = blockIdx.x % blocks_per_poly;
            double2 *base = V + poly * n + blk * pts;
            for (int i = t; i < pts; i += blockDim.x)
                sm[i] = base[i];
            __syncthreads();
            for (int s = local - 1; s >= 0; s--)
            {
                const int gap = 1 << s;
                const long long m = n >> (s + 1);
                for (int w = t; w < pts / 2; w += blockDim.x)
                {
                    const int g = w >> s, j = w & (gap - 1), xi = (g << (s + 1)) + j;
                    const long long gi = (blk << (local - s - 1)) + g;
                    const double2 u = sm[xi], v = cmul(sm[xi + gap], roots[m + gi]);
                    sm[xi] = cadd(u, v);
                    sm[xi + gap] = csubd(u, v);
                }
                __syncthreads();
            }
            for (int i = t; i < pts; i += blockDim.x)
                base[i] = sm[i];
        }
        // destination[i] = res[map[i]], i < n / 2 (ckks.h:782-785)
        __global__ void __launch_bounds__(256) ckks_gather_kernel(const double2 *__restrict__ V, const uint32_t *__restrict__ map, double2 *__restrict__ out,
                                                                  int logn, long long total)
        {
            const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; // over B * n / 2
            if (e >= total)
                return;
            const long long slots = 1ll << (logn - 1), b = e / slots, i = e % slots;
            out[e] = V[(b << logn) + map[i]];
        }

        unsigned grid_for(long long total, int threads)
        {
            return static_cast<unsigned>((total + threads - 1) / threads);
        }
    } // namespace

    Context::CkksEncoder &ckks_encoder(Context &c)
    {
        auto &enc = c.ckks;
        if (enc.ready)
            return enc;
        if (c.scheme != 2)
            throw std::invalid_argument("unsupported scheme");
        const sbh::CkksTables t = sbh::ckks_tables(c.n);
        const std::vector<uint32_t> map = sbh::batch_index_map(c.n); // matrix_reps_index_map_: the same map BatchEncoder uses
        cuda_check(cudaMalloc(&enc.d_roots, c.n * 16), "cudaMalloc(ckks roots)");
        cuda_check(cudaMalloc(&enc.d_inv_roots, c.n * 16), "cudaMalloc(ckks inverse roots)");
        cuda_check(cudaMalloc(&enc.d_map, c.n * sizeof(uint32_t)), "cudaMalloc(ckks index map)");
        cuda_check(cudaMalloc(&enc.d_stat, 16), "cudaMalloc(ckks status)");
        cuda_check(cudaMemcpy(enc.d_roots, t.roots.data(), c.n * 16, cudaMemcpyHostToDevice), "upload ckks roots");
        cuda_check(cudaMemcpy(enc.d_inv_roots, t.inv_roots.data(), c.n * 16, cudaMemcpyHostToDevice), "upload ckks inverse roots");
        cuda_check(cudaMemcpy(enc.d_map, map.data(), c.n * sizeof(uint32_t), cudaMemcpyHostToDevice), "upload ckks index map");
        c.table_bytes += c.n * 36;
        enc.ready = true;
        return enc;
    }

    static Context::CkksLevel &ckks_level(Context &c, size_t L)
    {
        auto it = c.ckks_levels.find(L);
        if (it != c.ckks_levels.end())
            return it->second;
        const sbh::CkksLevelHost h = sbh::ckks_level(c.q.data(), L);
        Context::CkksLevel lv;
        lv.total_bits = h.total_bits;
        std::vector<u64> big;
        big.insert(big.end(), h.Q.begin(), h.Q.end());
        big.insert(big.end(), h.threshold.begin(), h.threshold.end());
        big.insert(big.end(), h.punctured.begin(), h.punctured.end());
        std::vector<Tw> invp(L);
        for (size_t j = 0; j < L; j++)
            invp[j] = Tw{ h.inv_punctured[j].w, h.inv_punctured[j].wq };
        cuda_check(cudaMalloc(&lv.d_big, big.size() * sizeof(u64)), "cudaMalloc(ckks level)");
        cuda_check(cudaMalloc(&lv.d_invp, L * sizeof(Tw)), "cudaMalloc(ckks level)");
        cuda_check(cudaMemcpy(lv.d_big, big.data(), big.size() * sizeof(u64), cudaMemcpyHostToDevice), "upload ckks level");
        cuda_check(cudaMemcpy(lv.d_invp, invp.data(), L * sizeof(Tw), cudaMemcpyHostToDevice), "upload ckks level");
        return c.ckks_levels.emplace(L, lv).first->second;
    }

    int ckks_total_bits(Context &c, size_t L)
    {
        return ckks_level(c, L).total_bits;
    }

    // values: device, [B][count] complex (re, im) pairs or reals; plain: device, [B][L][n] NTT form.  Synchronises the stream once
    // (the reference's "encoded values are too large" check needs the largest coefficient before the plaintext is produced).
    void op_ckks_encode(Context &c, size_t L, size_t B, const double *values, size_t count, bool is_complex, double scale, u64 *plain,
                        cudaStream_t st)
    {
        auto &enc = ckks_encoder(c);
        if (L < 1 || L > c.k)
            throw std::invalid_argument("parms_id is not valid for encryption parameters");
        if (count > c.n / 2)
            throw std::invalid_argument("values_size is too large");
        if (count && !values)
            throw std::invalid_argument("values cannot be null");
        const int total_bits = ckks_level(c, L).total_bits;
        // ckks.h:494-499
        if (!std::isnormal(scale) || scale <= 0 || (static_cast<int>(std::log2(scale)) + 1 >= total_bits))
            throw std::invalid_argument("scale out of bounds");
        if (!B)
            return;
        const long long n = static_cast<long long>(c.n), logn = c.logn;
        double2 *V = static_cast<double2 *>(c.ensure_aux(B * c.n * sizeof(double2)));
        cuda_check(cudaMemsetAsync(V, 0, B * c.n * sizeof(double2), st), "memset");
        cuda_check(cudaMemsetAsync(enc.d_stat, 0, 16, st), "memset");
        unsigned long long *d_max = static_cast<unsigned long long *>(enc.d_stat);
        int *d_flags = reinterpret_cast<int *>(d_max + 1);
        if (count)
        {
            const long long total = static_cast<long long>(B) * static_cast<long long>(count);
            c.stats.begin("ckks_scatter", 0, 48.0 * total, st);
            ckks_scatter_kernel<<<grid_for(total, 256), 256, 0, st>>>(values, is_complex ? 1 : 0, static_cast<long long>(count), enc.d_map, V,
                                                                      static_cast<int>(logn), total, d_flags);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "ckks_scatter_kernel");
        }
        const double fix = scale / static_cast<double>(c.n); // ckks.h:522
        const int local = static_cast<int>(std::min<long long>(logn, kFftLocalLog));
        c.stats.begin("ckks_ifft_local", 0, 32.0 * B * n, st);
        ckks_ifft_local_kernel<<<static_cast<unsigned>(B * (c.n >> local)), std::min(1024, 1 << (local > 0 ? local - 1 : 0)),
                                 (size_t(1) << local) * sizeof(double2), st>>>(V, static_cast<const double2 *>(enc.d_inv_roots), static_cast<int>(logn), local, fix);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_ifft_local_kernel");
        const long long bflies = static_cast<long long>(B) * (n / 2);
        for (int s = local; s < logn; s++)
        {
            c.stats.begin("ckks_ifft_stage", s, 32.0 * B * n, st);
            if (s + 1 < logn)
                ckks_ifft_stage_kernel<false><<<grid_for(bflies, 256), 256, 0, st>>>(V, static_cast<const double2 *>(enc.d_inv_roots), static_cast<int>(logn), s, fix, bflies);
            else
                ckks_ifft_stage_kernel<true><<<grid_for(bflies, 256), 256, 0, st>>>(V, static_cast<const double2 *>(enc.d_inv_roots), static_cast<int>(logn), s, fix, bflies);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "ckks_ifft_stage_kernel");
        }
        const long long total = static_cast<long long>(B) * n;
        c.stats.begin("ckks_max", 0, 16.0 * total, st);
        ckks_max_kernel<<<grid_for(total, 256), 256, 0, st>>>(V, total, d_max, d_flags);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_max_kernel");
        struct
        {
            unsigned long long max_bits;
            int flags, pad;
        } stat;
        cuda_check(cudaMemcpyAsync(&stat, enc.d_stat, 16, cudaMemcpyDeviceToHost, st), "status D2H");
        cuda_check(cudaStreamSynchronize(st), "synchronize");
        if (stat.flags & 1)
            throw std::invalid_argument("values must be finite");
        double max_coeff;
        std::memcpy(&max_coeff, &stat.max_bits, sizeof(double));
        // ckks.h:533-553
        if ((stat.flags & 2) || !std::isfinite(max_coeff))
            throw std::invalid_argument("encoded values are too large");
        const int max_coeff_bit_count = static_cast<int>(std::ceil(std::log2(std::max(max_coeff, 1.0)))) + 1;
        if (max_coeff_bit_count >= total_bits)
            throw std::invalid_argument("encoded values are too large");
        c.stats.begin("ckks_reduce", 0, (16.0 + 8.0 * L) * total, st);
        ckks_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(V, plain, c.d_primes, static_cast<int>(logn), static_cast<int>(L), total);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_reduce_kernel");
        op_ntt(c, false, L, 1, B, plain, st);
    }

    // plain: device [B][L][n] NTT form, scale = Plaintext::scale(); values out: device [B][n/2] complex (re, im)
    void op_ckks_decode(Context &c, size_t L, size_t B, const u64 *plain, double scale, double *values, cudaStream_t st)
    {
        auto &enc = ckks_encoder(c);
        if (L < 1 || L > c.k || L > static_cast<size_t>(kMaxWords))
            throw std::invalid_argument("plain is not valid for encryption parameters");
        auto &lv = ckks_level(c, L);
        // ckks.h:712-716
        if (!std::isnormal(scale) || scale <= 0 || (static_cast<int>(std::log2(scale)) >= lv.total_bits))
            throw std::invalid_argument("scale out of bounds");
        if (!B)
            return;
        const long long n = static_cast<long long>(c.n), logn = c.logn, total = static_cast<long long>(B) * n;
        const size_t coef_bytes = B * L * c.n * sizeof(u64);
        unsigned char *arena = static_cast<unsigned char *>(c.ensure_aux(coef_bytes + B * c.n * sizeof(double2)));
        u64 *coef = reinterpret_cast<u64 *>(arena);
        double2 *V = reinterpret_cast<double2 *>(arena + coef_bytes);
        cuda_check(cudaMemcpyAsync(coef, plain, coef_bytes, cudaMemcpyDeviceToDevice, st), "copy");
        op_ntt(c, true, L, 1, B, coef, st);
        const double inv_scale = 1.0 / scale;
        c.stats.begin("ckks_compose", 0, (8.0 * L + 16.0) * total, st);
        ckks_compose_kernel<<<grid_for(total, 128), 128, 0, st>>>(coef, lv.d_big, lv.d_invp, c.d_primes, V, static_cast<int>(logn), static_cast<int>(L),
                                                                  inv_scale, total);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_compose_kernel");
        const int local = static_cast<int>(std::min<long long>(logn, kFftLocalLog));
        const long long bflies = static_cast<long long>(B) * (n / 2);
        for (int s = static_cast<int>(logn) - 1; s >= local; s--)
        {
            c.stats.begin("ckks_fft_stage", s, 32.0 * B * n, st);
            ckks_fft_stage_kernel<<<grid_for(bflies, 256), 256, 0, st>>>(V, static_cast<const double2 *>(enc.d_roots), static_cast<int>(logn), s, bflies);
            c.stats.end(st);
            cuda_check(cudaGetLastError(), "ckks_fft_stage_kernel");
        }
        c.stats.begin("ckks_fft_local", 0, 32.0 * B * n, st);
        ckks_fft_local_kernel<<<static_cast<unsigned>(B * (c.n >> local)), std::min(1024, 1 << (local > 0 ? local - 1 : 0)),
                                (size_t(1) << local) * sizeof(double2), st>>>(V, static_cast<const double2 *>(enc.d_roots), static_cast<int>(logn), local);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_fft_local_kernel");
        c.stats.begin("ckks_gather", 0, 16.0 * total, st);
        ckks_gather_kernel<<<grid_for(total / 2, 256), 256, 0, st>>>(V, enc.d_map, reinterpret_cast<double2 *>(values), static_cast<int>(logn), total / 2);
        c.stats.end(st);
        cuda_check(cudaGetLastError(), "ckks_gather_kernel");
    }
} // namespace sb
