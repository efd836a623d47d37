// tools/bfly2.cu -- candidate formulations of the hot arithmetic (forward / inverse butterflies, key multiply-accumulate)
// measured in isolation at the occupancies the real kernels run at, each checked against plain __int128 arithmetic.
// Cost model behind the candidates (tools/issue_microbench.cu on B200): IMAD.WIDE occupies the multiply pipe for 4 clocks per
// warp, IMAD lo 2, every integer-ALU instruction (IADD3, LOP3, SHF, SEL, ISETP) 2 on its own pipe, with partial overlap
// (time ~ multiply-pipe time + 0.45 x ALU time).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bfly2 bfly2.cu
#include "../seal_b200/csrc/sb_device.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned __int128 u128;
#define ROUNDS 256

// ---------------------------------------------------------------- candidates ----
// round-1 quotient estimate: carry of the two high halves materialised with add.cc / addc, then one fused wide multiply-add
__device__ __forceinline__ u64 approx_mulhi_r1(u64 y, u64 wq)
{
    unsigned y0, y1, wq0, wq1, alo, ahi, blo, bhi, slo, shi;
    unpack64(y, y0, y1);
    unpack64(wq, wq0, wq1);
    u64 a, b, T;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(y1), "r"(wq0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(y0), "r"(wq1));
    unpack64(a, alo, ahi);
    unpack64(b, blo, bhi);
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(slo), "=r"(shi) : "r"(ahi), "r"(bhi));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(T) : "r"(y1), "r"(wq1), "l"(pack64(slo, shi)));
    return T;
}
template <bool FAST>
__device__ __forceinline__ void ct_bfly_r1(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = mullo_combine(y, w.w, approx_mulhi_r1(y, w.wq), P.nq);
    u64 u = FAST ? x : csub(x, P.q4);
    x = u + v + P.zero;
    y = u - v + P.q4;
}
__device__ __forceinline__ void gs_bfly_r1(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 u = x, v = y;
    x = csub(u + v, P.q4);
    u64 t = u - v + P.q4;
    y = mullo_combine(t, w.w, approx_mulhi_r1(t, w.wq), P.nq);
}
// quotient high halves through IMAD.HI (no 64-bit temporaries for the cross products)
__device__ __forceinline__ u64 approx_mulhi_hi(u64 y, u64 wq)
{
    unsigned y0, y1, wq0, wq1;
    unpack64(y, y0, y1);
    unpack64(wq, wq0, wq1);
    const unsigned ahi = __umulhi(y1, wq0), bhi = __umulhi(y0, wq1);
    return static_cast<u64>(y1) * wq1 + static_cast<u64>(ahi) + static_cast<u64>(bhi);
}
template <bool FAST>
__device__ __forceinline__ void ct_bfly_hi(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = mullo_combine(y, w.w, approx_mulhi_hi(y, w.wq), P.nq);
    u64 u = FAST ? x : csub(x, P.q4);
    x = u + v + P.zero;
    y = u - v + P.q4;
}

// ---- multiply-accumulate candidates (the shipped one is mac128 of sb_device.cuh) ----
// 28-bit limbs: for primes q = 2^bits - dsol below 2^56 the transformed digit is folded below 2^56 (Solinas: x = (x mod 2^bits) +
// (x >> bits) * dsol), split into limbs a = a1*2^28 + a0 and multiplied with the key word kept in the same limb form into three
// plain 64-bit column sums without carries (Karatsuba: 3 wide multiplies per product; schoolbook: 4).  Measured (this file, and the
// fused key-switch kernel built both ways, profiles/r02_experiments.md): fewer multiply-pipe clocks, but three 64-bit sums per
// accumulator instead of two -- with half of the sums in shared memory the real kernel got 6 % slower, so mac128 stayed.
__device__ __forceinline__ u64 fold_solinas(u64 a, unsigned bits, unsigned dsol)
{
    const u64 low = a & ((1ull << bits) - 1ull);
    return low + static_cast<u64>(static_cast<unsigned>(a >> bits)) * dsol;
}
struct Acc3
{
    u64 s0, s1, s2;
};
__device__ __forceinline__ void mac_limb28(Acc3 &s, unsigned a0, unsigned a1, unsigned as, u64 ke)
{
    unsigned k0, k1;
    unpack64(ke, k0, k1);
    const unsigned ks = k0 + k1;
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s0) : "r"(a0), "r"(k0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s2) : "r"(a1), "r"(k1));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s1) : "r"(as), "r"(ks));
}
__device__ __forceinline__ void acc3_value(const Acc3 &s, u64 &lo, u64 &hi)
{
    const u64 mid = s.s1 - s.s0 - s.s2; // sum of a0*k1 + a1*k0
    u64 l = s.s0, h = 0, t;
    t = mid << 28;
    l += t, h += (mid >> 36) + (l < t);
    t = s.s2 << 56;
    l += t, h += (s.s2 >> 8) + (l < t);
    lo = l, hi = h;
}
// schoolbook on 28-bit limbs: four wide multiplies, three column sums
__device__ __forceinline__ void mac_limb28_school(Acc3 &s, unsigned a0, unsigned a1, u64 klimbs)
{
    unsigned k0, k1;
    unpack64(klimbs, k0, k1);
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s0) : "r"(a0), "r"(k0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s1) : "r"(a0), "r"(k1));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s1) : "r"(a1), "r"(k0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s.s2) : "r"(a1), "r"(k1));
}
__device__ __forceinline__ void acc3_school_value(const Acc3 &s, u64 &lo, u64 &hi)
{
    Acc3 t{ s.s0, s.s1 + s.s0 + s.s2, s.s2 }; // acc3_value subtracts s0 + s2 from the middle column (Karatsuba form)
    acc3_value(t, lo, hi);
}

// ---------------------------------------------------------------- kernels ----
// KIND: 0 ct_bfly<true> (shipped), 1 round-1 quotient, 2 IMAD.HI quotient, 3 ct_bfly<false> (shipped guarded), 4 guarded round-1, 5 gs shipped, 6 gs round-1
template <int KIND, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) kb(u64 *d, const PrimeDev *pp, const Tw *tws, int check)
{
    __shared__ Tw ts[256];
    const PrimeDev P = *pp;
    for (int i = threadIdx.x; i < 256; i += THREADS)
        ts[i] = tws[i];
    __syncthreads();
    u64 a[8];
    for (int j = 0; j < 8; j++)
        a[j] = d[(blockIdx.x * 8 + j) * THREADS + threadIdx.x];
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < (check ? 1 : ROUNDS); r++)
    {
        const Tw *t = ts + ((r * 7 + lane) & 127); // per-lane twiddles from shared memory, as in the transform kernels
#define BF(X, Y, W)                                         \
    if (KIND == 0) ct_bfly<true>(X, Y, W, P);               \
    else if (KIND == 1) ct_bfly_r1<true>(X, Y, W, P);       \
    else if (KIND == 2) ct_bfly_hi<true>(X, Y, W, P);       \
    else if (KIND == 3) ct_bfly<false>(X, Y, W, P);         \
    else if (KIND == 4) ct_bfly_r1<false>(X, Y, W, P);      \
    else if (KIND == 5) gs_bfly(X, Y, W, P);                \
    else gs_bfly_r1(X, Y, W, P);
#pragma unroll
        for (int j = 0; j < 4; j++) { BF(a[j], a[j + 4], t[0]) }
        BF(a[0], a[2], t[1]) BF(a[1], a[3], t[1]) BF(a[4], a[6], t[2]) BF(a[5], a[7], t[2])
#pragma unroll
        for (int p = 0; p < 4; p++) { BF(a[2 * p], a[2 * p + 1], t[3 + p]) }
        if (KIND <= 2 && (r & 3) == 3)
        {
#pragma unroll
            for (int j = 0; j < 8; j++)
                a[j] = barrett_lazy4(a[j], P.ratio_hi, P.nq);
        }
    }
    for (int j = 0; j < 8; j++)
        d[(blockIdx.x * 8 + j) * THREADS + threadIdx.x] = a[j];
}

// multiply-accumulate of 8 coefficients x 2 key components x DIG digits per thread; values arrive lazily grown (< 2^62)
// KIND 0: mac128 (both components in registers); 1: fold + Karatsuba limbs (both in registers); 2: mac128, component 1 in shared
// memory (as the fused kernel does); 3: Karatsuba limbs, component 1 in shared memory; 4 / 5: schoolbook limbs, registers / shared
#define MACL(S, J, KW)                                                       \
    do                                                                       \
    {                                                                        \
        if (KIND == 4 || KIND == 5) mac_limb28_school(S, a0[J], a1[J], KW);  \
        else mac_limb28(S, a0[J], a1[J], as[J], KW);                         \
    } while (0)
#define ACCV(S, LO, HI)                                                      \
    do                                                                       \
    {                                                                        \
        if (KIND == 4 || KIND == 5) acc3_school_value(S, LO, HI);            \
        else acc3_value(S, LO, HI);                                          \
    } while (0)
#define DIG 31
template <int KIND>
__global__ void __launch_bounds__(256, 2) km(const u64 *__restrict__ vals, const u64 *__restrict__ key, u64 *out, const PrimeDev *pp, int b, unsigned dsol)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const PrimeDev P = *pp;
    const int tid = threadIdx.x;
    u64 lo0[8], hi0[8], lo1[8], hi1[8];
    Acc3 s0[8], s1[8];
    ulonglong2(*acc1)[256] = reinterpret_cast<ulonglong2(*)[256]>(sm);
    u64(*acc3)[3][256] = reinterpret_cast<u64(*)[3][256]>(sm);
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        lo0[j] = hi0[j] = lo1[j] = hi1[j] = 0;
        s0[j] = Acc3{ 0, 0, 0 }, s1[j] = Acc3{ 0, 0, 0 };
        if (KIND == 2)
            acc1[j][tid] = make_ulonglong2(0, 0);
        if (KIND == 3 || KIND == 5)
            acc3[j][0][tid] = acc3[j][1][tid] = acc3[j][2][tid] = 0;
    }
    for (int J = 0; J < DIG; J++)
    {
        const ulonglong2 *vp = reinterpret_cast<const ulonglong2 *>(vals + ((static_cast<long long>(blockIdx.x) * DIG + J) * 256 + tid) * 8);
        const ulonglong2 *k0 = reinterpret_cast<const ulonglong2 *>(key + ((static_cast<long long>(J) * 2 + 0) * 256 + tid) * 8);
        const ulonglong2 *k1 = reinterpret_cast<const ulonglong2 *>(key + ((static_cast<long long>(J) * 2 + 1) * 256 + tid) * 8);
        u64 a[8];
#pragma unroll
        for (int h = 0; h < 4; h++)
        {
            ulonglong2 v = vp[h];
            a[2 * h] = v.x, a[2 * h + 1] = v.y;
        }
        if (KIND == 0 || KIND == 2)
        {
#pragma unroll
            for (int h = 0; h < 4; h++)
            {
                ulonglong2 v = __ldg(k0 + h);
                mac128(lo0[2 * h], hi0[2 * h], a[2 * h], v.x);
                mac128(lo0[2 * h + 1], hi0[2 * h + 1], a[2 * h + 1], v.y);
            }
#pragma unroll
            for (int h = 0; h < 4; h++)
            {
                ulonglong2 v = __ldg(k1 + h);
                if (KIND == 0)
                {
                    mac128(lo1[2 * h], hi1[2 * h], a[2 * h], v.x);
                    mac128(lo1[2 * h + 1], hi1[2 * h + 1], a[2 * h + 1], v.y);
                }
                else
                {
                    ulonglong2 t0 = acc1[2 * h][tid], t1 = acc1[2 * h + 1][tid];
                    mac128(t0.x, t0.y, a[2 * h], v.x);
                    mac128(t1.x, t1.y, a[2 * h + 1], v.y);
                    acc1[2 * h][tid] = t0, acc1[2 * h + 1][tid] = t1;
                }
            }
        }
        else
        {
            unsigned a0[8], a1[8], as[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
            {
                u64 f = fold_solinas(a[j], b, dsol);
                a0[j] = static_cast<unsigned>(f) & 0x0FFFFFFFu, a1[j] = static_cast<unsigned>(f >> 28), as[j] = a0[j] + a1[j];
            }
#pragma unroll
            for (int h = 0; h < 4; h++)
            {
                ulonglong2 v = __ldg(k0 + h);
                MACL(s0[2 * h], 2 * h, v.x);
                MACL(s0[2 * h + 1], 2 * h + 1, v.y);
            }
#pragma unroll
            for (int h = 0; h < 4; h++)
            {
                ulonglong2 v = __ldg(k1 + h);
                if (KIND == 1 || KIND == 4)
                {
                    MACL(s1[2 * h], 2 * h, v.x);
                    MACL(s1[2 * h + 1], 2 * h + 1, v.y);
                }
                else
                {
#pragma unroll
                    for (int e = 0; e < 2; e++)
                    {
                        Acc3 t{ acc3[2 * h + e][0][tid], acc3[2 * h + e][1][tid], acc3[2 * h + e][2][tid] };
                        MACL(t, 2 * h + e, e ? v.y : v.x);
                        acc3[2 * h + e][0][tid] = t.s0, acc3[2 * h + e][1][tid] = t.s1, acc3[2 * h + e][2][tid] = t.s2;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        u64 l0, h0, l1, h1;
        if (KIND == 0 || KIND == 2)
        {
            l0 = lo0[j], h0 = hi0[j];
            if (KIND == 0)
                l1 = lo1[j], h1 = hi1[j];
            else
                l1 = acc1[j][tid].x, h1 = acc1[j][tid].y;
        }
        else
        {
            ACCV(s0[j], l0, h0);
            if (KIND == 1 || KIND == 4)
                ACCV(s1[j], l1, h1);
            else
                ACCV((Acc3{ acc3[j][0][tid], acc3[j][1][tid], acc3[j][2][tid] }), l1, h1);
        }
        out[((static_cast<long long>(blockIdx.x) * 2 + 0) * 256 + tid) * 8 + j] = barrett128(l0, h0, P.q, P.ratio_lo, P.ratio_hi);
        out[((static_cast<long long>(blockIdx.x) * 2 + 1) * 256 + tid) * 8 + j] = barrett128(l1, h1, P.q, P.ratio_lo, P.ratio_hi);
    }
}

// ---------------------------------------------------------------- host ----
static u64 g_q = 36028797017456641ull; // 2^55 - 12*2^17 + 1
static double time_ms(cudaEvent_t e0, cudaEvent_t e1)
{
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
template <int KIND, int THREADS, int MINB>
static void run_b(const char *name, u64 *d, PrimeDev *dp, Tw *dt, const std::vector<u64> &h0, const std::vector<Tw> &ht)
{
    const int blocks = 148 * MINB;
    const size_t words = static_cast<size_t>(blocks) * 8 * THREADS;
    // correctness: one round against plain 128-bit arithmetic
    cudaMemcpy(d, h0.data(), words * 8, cudaMemcpyHostToDevice);
    kb<KIND, THREADS, MINB><<<blocks, THREADS>>>(d, dp, dt, 1);
    std::vector<u64> got(words);
    cudaMemcpy(got.data(), d, words * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    const u64 q = g_q;
    auto mm = [q](u64 a, u64 b) { return static_cast<u64>(static_cast<u128>(a % q) * (b % q) % q); };
    for (int blk = 0; blk < 2; blk++)
        for (int tid = 0; tid < THREADS; tid++)
        {
            u64 a[8];
            for (int j = 0; j < 8; j++)
                a[j] = h0[(static_cast<size_t>(blk) * 8 + j) * THREADS + tid] % q;
            const Tw *t = ht.data() + ((tid & 31) & 127);
            auto bf = [&](u64 &x, u64 &y, Tw w) {
                if (KIND <= 4)
                {
                    u64 v = mm(y, w.w), u = x;
                    x = (u + v) % q, y = (u + q - v) % q;
                }
                else
                {
                    u64 u = x, v = y;
                    x = (u + v) % q, y = mm((u + q - v) % q, w.w);
                }
            };
            for (int j = 0; j < 4; j++)
                bf(a[j], a[j + 4], t[0]);
            bf(a[0], a[2], t[1]), bf(a[1], a[3], t[1]), bf(a[4], a[6], t[2]), bf(a[5], a[7], t[2]);
            for (int p = 0; p < 4; p++)
                bf(a[2 * p], a[2 * p + 1], t[3 + p]);
            for (int j = 0; j < 8; j++)
                if (got[(static_cast<size_t>(blk) * 8 + j) * THREADS + tid] % q != a[j])
                    bad++;
        }
    cudaMemcpy(d, h0.data(), words * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    kb<KIND, THREADS, MINB><<<blocks, THREADS>>>(d, dp, dt, 0);
    cudaEventRecord(e0);
    kb<KIND, THREADS, MINB><<<blocks, THREADS>>>(d, dp, dt, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kb<KIND, THREADS, MINB>);
    const double warp_bf = static_cast<double>(blocks) * (THREADS / 32) * ROUNDS * 12.0;
    const double ms = time_ms(e0, e1);
    printf("%-34s %3d thr x %d CTA/SM  regs %3d  %7.3f ms  %6.2f clk per warp-butterfly per SMSP  %s\n", name, THREADS, MINB, fa.numRegs, ms,
           ms * 1e-3 * khz * 1e3 / (warp_bf / (148.0 * 4)), bad ? "MISMATCH" : "ok");
}
template <int KIND>
static void run_m(const char *name, const PrimeDev &P, PrimeDev *dp)
{
    const int blocks = 148 * 2 * 4;
    const size_t nv = static_cast<size_t>(blocks) * DIG * 256 * 8, nk = static_cast<size_t>(DIG) * 2 * 256 * 8, no = static_cast<size_t>(blocks) * 2 * 256 * 8;
    std::vector<u64> hv(nv), hk(nk), hkl(nk);
    u64 x = 88172645463325252ull;
    auto rnd = [&x]() { x ^= x << 13, x ^= x >> 7, x ^= x << 17; return x; };
    for (auto &v : hv)
        v = rnd() >> 2; // lazily grown transform output: anything below 2^62
    for (size_t i = 0; i < nk; i++)
    {
        hk[i] = rnd() % P.q;
        hkl[i] = (hk[i] & 0x0FFFFFFFull) | ((hk[i] >> 28) << 32);
    }
    u64 *dv, *dk, *dout;
    cudaMalloc(&dv, nv * 8), cudaMalloc(&dk, nk * 8), cudaMalloc(&dout, no * 8);
    cudaMemcpy(dv, hv.data(), nv * 8, cudaMemcpyHostToDevice);
    const bool limbs = KIND != 0 && KIND != 2;
    cudaMemcpy(dk, (limbs ? hkl : hk).data(), nk * 8, cudaMemcpyHostToDevice);
    const int b = 55;
    const unsigned dsol = static_cast<unsigned>((1ull << b) - P.q);
    const size_t smem = KIND == 2 ? 8 * 256 * 16 : ((KIND == 3 || KIND == 5) ? 8 * 3 * 256 * 8 : 0);
    cudaFuncSetAttribute(km<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    km<KIND><<<blocks, 256, smem>>>(dv, dk, dout, dp, b, dsol);
    cudaEventRecord(e0);
    km<KIND><<<blocks, 256, smem>>>(dv, dk, dout, dp, b, dsol);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    std::vector<u64> got(2 * 256 * 8);
    cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int c = 0; c < 2; c++)
        for (int tid = 0; tid < 256; tid += 37)
            for (int j = 0; j < 8; j++)
            {
                u128 s = 0;
                for (int J = 0; J < DIG; J++)
                    s = (s + static_cast<u128>(hv[((static_cast<size_t>(0) * DIG + J) * 256 + tid) * 8 + j] % P.q) *
                                 hk[((static_cast<size_t>(J) * 2 + c) * 256 + tid) * 8 + j]) % P.q;
                if (got[(static_cast<size_t>(c) * 256 + tid) * 8 + j] % P.q != static_cast<u64>(s))
                    bad++;
            }
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, km<KIND>);
    const double ms = time_ms(e0, e1), warp_mac = static_cast<double>(blocks) * 8 * DIG * 16.0;
    printf("%-46s regs %3d  %7.3f ms  %6.2f clk per warp-MAC per SMSP (incl. loads)  %s\n", name, fa.numRegs, ms,
           ms * 1e-3 * khz * 1e3 / (warp_mac / (148.0 * 4)), bad ? "MISMATCH" : "ok");
    cudaFree(dv), cudaFree(dk), cudaFree(dout);
}
int main()
{
    const u64 q = g_q;
    PrimeDev P{};
    P.q = q, P.q2 = 2 * q, P.q4 = 4 * q, P.nq = 0ull - q;
    u128 all = ~static_cast<u128>(0);
    P.ratio_lo = static_cast<u64>(all / q), P.ratio_hi = static_cast<u64>((all / q) >> 64);
    std::vector<Tw> ht(256);
    u64 x = 0x9E3779B97F4A7C15ull;
    auto rnd = [&x]() { x ^= x << 13, x ^= x >> 7, x ^= x << 17; return x; };
    for (auto &t : ht)
    {
        t.w = rnd() % q;
        t.wq = static_cast<u64>((static_cast<u128>(t.w) << 64) / q);
    }
    const size_t maxw = static_cast<size_t>(148) * 4 * 8 * 512;
    std::vector<u64> h0(maxw);
    for (auto &v : h0)
        v = rnd() % q;
    u64 *d;
    PrimeDev *dp;
    Tw *dt;
    cudaMalloc(&d, maxw * 8), cudaMalloc(&dp, sizeof(P)), cudaMalloc(&dt, ht.size() * sizeof(Tw));
    cudaMemcpy(dp, &P, sizeof(P), cudaMemcpyHostToDevice);
    cudaMemcpy(dt, ht.data(), ht.size() * sizeof(Tw), cudaMemcpyHostToDevice);
    run_b<0, 256, 2>("fwd guard-free, shipped", d, dp, dt, h0, ht);
    run_b<1, 256, 2>("fwd guard-free, round-1 quotient", d, dp, dt, h0, ht);
    run_b<2, 256, 2>("fwd guard-free, IMAD.HI quotient", d, dp, dt, h0, ht);
    run_b<3, 256, 2>("fwd guarded, shipped", d, dp, dt, h0, ht);
    run_b<4, 256, 2>("fwd guarded, round-1 quotient", d, dp, dt, h0, ht);
    run_b<5, 256, 2>("inv, shipped", d, dp, dt, h0, ht);
    run_b<6, 256, 2>("inv, round-1 quotient", d, dp, dt, h0, ht);
    run_b<0, 512, 3>("fwd guard-free, shipped", d, dp, dt, h0, ht);
    run_b<1, 512, 3>("fwd guard-free, round-1 quotient", d, dp, dt, h0, ht);
    run_b<2, 512, 3>("fwd guard-free, IMAD.HI quotient", d, dp, dt, h0, ht);
    run_b<0, 512, 2>("fwd guard-free, shipped", d, dp, dt, h0, ht);
    run_b<1, 512, 2>("fwd guard-free, round-1 quotient", d, dp, dt, h0, ht);
    run_b<0, 256, 4>("fwd guard-free, shipped", d, dp, dt, h0, ht);
    run_b<1, 256, 4>("fwd guard-free, round-1 quotient", d, dp, dt, h0, ht);
    run_b<5, 256, 4>("inv, shipped", d, dp, dt, h0, ht);
    run_b<6, 256, 4>("inv, round-1 quotient", d, dp, dt, h0, ht);
    run_m<0>("MAC 128-bit carry chain, registers", P, dp);
    run_m<1>("MAC fold + Karatsuba limbs, registers", P, dp);
    run_m<4>("MAC fold + schoolbook limbs, registers", P, dp);
    run_m<2>("MAC 128-bit carry chain, comp. 1 in smem", P, dp);
    run_m<3>("MAC fold + Karatsuba limbs, comp. 1 in smem", P, dp);
    run_m<5>("MAC fold + schoolbook limbs, comp. 1 in smem", P, dp);
    return 0;
}
