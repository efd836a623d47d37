// tools/bfly_microbench.cu -- cycles per butterfly of the device arithmetic in isolation (registers only).
#include "../seal_b200/csrc/sb_device.cuh"
#include <cstdio>
#define ROUNDS 512
// ---------------- candidate implementations ----------------
__device__ __forceinline__ void bf_exact_guard(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{   // v1 of this repo: exact Shoup + guard, values < 4q
    u64 u = csub(x, P.q2);
    u64 v = mul_shoup_lazy(y, w, P.q);
    x = u + v;
    y = u - v + P.q2;
}
__device__ __forceinline__ u64 approx_mulhi_w(u64 y, u64 wq)
{   // all three partial products as mul.wide (no IMAD.HI)
    unsigned y0 = (unsigned)y, y1 = (unsigned)(y >> 32), wq0 = (unsigned)wq, wq1 = (unsigned)(wq >> 32);
    u64 a, b, T;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(y1), "r"(wq0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(y0), "r"(wq1));
    u64 s = (a >> 32) + (b >> 32);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(T) : "r"(y1), "r"(wq1), "l"(s));
    return T;
}
__device__ __forceinline__ u64 mullo_combine_a(u64 y, u64 w, u64 T, u64 nq)
{   // everything through mad chains
    unsigned y0 = (unsigned)y, y1 = (unsigned)(y >> 32), w0 = (unsigned)w, w1 = (unsigned)(w >> 32);
    unsigned T0 = (unsigned)T, T1 = (unsigned)(T >> 32), n0 = (unsigned)nq, n1 = (unsigned)(nq >> 32);
    u64 acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(T0), "r"(n0));
    unsigned lo = (unsigned)acc, hi = (unsigned)(acc >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T1), "r"(n0));
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ void bf_fast_asm(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = mullo_combine_a(y, w.w, approx_mulhi_w(y, w.wq), P.nq);
    u64 u = x;
    x = u + v;
    y = u - v + P.q4;
}
__device__ __forceinline__ void bf_fast_exact(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{   // exact quotient (compiler's mul.hi.u64), remainder via 64-bit mad.lo with 2^64-q: result < 2q, growth 2q/stage
    u64 T = __umul64hi(y, w.wq);
    u64 v = y * w.w + T * P.nq;
    u64 u = x;
    x = u + v;
    y = u - v + P.q2;
}
__device__ __forceinline__ void bf_fast_asm2(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{   // approximate quotient, remainder through plain 64-bit C multiplies (compiler picks the expansion)
    u64 T = approx_mulhi_w(y, w.wq);
    u64 v = y * w.w + T * P.nq;
    u64 u = x;
    x = u + v;
    y = u - v + P.q4;
}
__device__ __forceinline__ u64 pack(unsigned lo, unsigned hi){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, unsigned &lo, unsigned &hi){ asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ u64 lazy4_asm(u64 y, Tw w, u64 nq)
{
    unsigned y0, y1, wq0, wq1, w0, w1, n0, n1;
    unpack(y, y0, y1); unpack(w.wq, wq0, wq1); unpack(w.w, w0, w1); unpack(nq, n0, n1);
    u64 a, b, T, acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(y1), "r"(wq0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(y0), "r"(wq1));
    unsigned alo, ahi, blo, bhi; unpack(a, alo, ahi); unpack(b, blo, bhi);
    unsigned slo, shi;
    asm("add.cc.u32 %0, %2, %3; addc.u32 %1, 0, 0;" : "=r"(slo), "=r"(shi) : "r"(ahi), "r"(bhi));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(T) : "r"(y1), "r"(wq1), "l"(pack(slo, shi)));
    unsigned T0, T1; unpack(T, T0, T1);
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(T0), "r"(n0));
    unsigned lo, hi; unpack(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T1), "r"(n0));
    return pack(lo, hi);
}
__device__ __forceinline__ u64 lazy4_asm_b(u64 y, Tw w, u64 nq)
{   // variant: the four 32-bit cross products as two independent chains (shorter dependency chain)
    unsigned y0, y1, wq0, wq1, w0, w1, n0, n1;
    unpack(y, y0, y1); unpack(w.wq, wq0, wq1); unpack(w.w, w0, w1); unpack(nq, n0, n1);
    u64 a, b, T, acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(y1), "r"(wq0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(y0), "r"(wq1));
    unsigned c1 = y0 * w1;                       // independent of T: can issue while T is being formed
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c1) : "r"(y1), "r"(w0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
    unsigned alo, ahi, blo, bhi; unpack(a, alo, ahi); unpack(b, blo, bhi);
    unsigned slo, shi;
    asm("add.cc.u32 %0, %2, %3; addc.u32 %1, 0, 0;" : "=r"(slo), "=r"(shi) : "r"(ahi), "r"(bhi));
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(T) : "r"(y1), "r"(wq1), "l"(pack(slo, shi)));
    unsigned T0, T1; unpack(T, T0, T1);
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(T0), "r"(n0));
    unsigned c2 = T0 * n1;
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c2) : "r"(T1), "r"(n0));
    unsigned lo, hi; unpack(acc, lo, hi);
    hi = hi + c1 + c2;
    return pack(lo, hi);
}
__device__ __forceinline__ void bf_fast_asm4(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = lazy4_asm_b(y, w, P.nq);
    u64 u = x;
    x = u + v;
    y = u - v + P.q4;
}
__device__ __forceinline__ void bf_fast_asm3(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = lazy4_asm(y, w, P.nq);
    u64 u = x;
    x = u + v;
    y = u - v + P.q4;
}
__device__ __forceinline__ void bf_guard_asm3(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = lazy4_asm(y, w, P.nq);
    u64 u = csub(x, P.q4);
    x = u + v;
    y = u - v + P.q4;
}
template <int KIND>
__global__ void __launch_bounds__(256) k(u64 *d, const PrimeDev *pp, const Tw *tws)
{
    const PrimeDev P = *pp;
    u64 a[8];
    Tw t[7];
    for (int j = 0; j < 8; j++) a[j] = d[threadIdx.x + 256 * j] % P.q;
    for (int j = 0; j < 7; j++) t[j] = tws[j];
    for (int r = 0; r < ROUNDS; r++)
    {
#define BF(X, Y, W)                                              \
    if (KIND == 0) bf_exact_guard(X, Y, W, P);                   \
    else if (KIND == 1) ct_bfly<true>(X, Y, W, P);               \
    else if (KIND == 2) ct_bfly<false>(X, Y, W, P);              \
    else if (KIND == 3) bf_fast_asm(X, Y, W, P);                 \
    else if (KIND == 4) gs_bfly(X, Y, W, P);                     \
    else if (KIND == 5) bf_fast_exact(X, Y, W, P);               \
    else if (KIND == 6) bf_fast_asm2(X, Y, W, P);                \
    else if (KIND == 7) bf_fast_asm3(X, Y, W, P);                \
    else if (KIND == 8) bf_guard_asm3(X, Y, W, P);               \
    else if (KIND == 9) bf_fast_asm4(X, Y, W, P);
#pragma unroll
        for (int j = 0; j < 4; j++) { BF(a[j], a[j + 4], t[0]) }
        BF(a[0], a[2], t[1]) BF(a[1], a[3], t[1]) BF(a[4], a[6], t[2]) BF(a[5], a[7], t[2])
#pragma unroll
        for (int p = 0; p < 4; p++) { BF(a[2 * p], a[2 * p + 1], t[3 + p]) }
        if (KIND == 1 || KIND == 3 || KIND == 5 || KIND == 6 || KIND == 7 || KIND == 9)
        {   // keep FAST-mode values bounded the way a real kernel does once per 17 stages; here once per 12
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = (r & 7) ? a[j] : barrett_lazy4(a[j], P.ratio_hi, P.nq);
        }
    }
    u64 s = 0;
    for (int j = 0; j < 8; j++) s ^= a[j];
    d[blockIdx.x * 256 + threadIdx.x] = s;
}
static int g_blocks_per_sm = 4;
template <int KIND>
void run(const char *name, u64 *d, PrimeDev *dp, Tw *dt)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int blocks = 148 * g_blocks_per_sm;
    k<KIND><<<blocks, 256>>>(d, dp, dt);
    cudaEventRecord(e0);
    k<KIND><<<blocks, 256>>>(d, dp, dt);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_bf = (double)blocks * 8 * ROUNDS * 12.0;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-36s %8.3f ms   %6.2f cycles per warp-butterfly per SMSP\n", name, ms, cycles / (warp_bf / (148.0 * 4)));
}
int main()
{
    u64 q = 36028797017456641ull; // 55-bit prime = 1 mod 2^17
    PrimeDev P{};
    P.q = q, P.q2 = 2 * q, P.q4 = 4 * q, P.nq = 0ull - q;
    unsigned __int128 all = ~(unsigned __int128)0;
    P.ratio_lo = (u64)(all / q), P.ratio_hi = (u64)((all / q) >> 64);
    Tw t[7];
    for (int j = 0; j < 7; j++)
    {
        t[j].w = (q / 7) * (j + 1) + 12345;
        t[j].wq = (u64)((((unsigned __int128)t[j].w) << 64) / q);
    }
    u64 *d;
    PrimeDev *dp;
    Tw *dt;
    cudaMalloc(&d, 148 * 8 * 256 * 8 * 8);
    cudaMemset(d, 7, 148 * 8 * 256 * 8 * 8);
    cudaMalloc(&dp, sizeof(P));
    cudaMalloc(&dt, sizeof(t));
    cudaMemcpy(dp, &P, sizeof(P), cudaMemcpyHostToDevice);
    cudaMemcpy(dt, t, sizeof(t), cudaMemcpyHostToDevice);
    run<0>("v1 exact Shoup + guard", d, dp, dt);
    run<1>("FAST lazy4 (C), + barrett/96 stages", d, dp, dt);
    run<2>("guarded lazy4 (C)", d, dp, dt);
    run<3>("FAST lazy4 (asm wide), + barrett/96", d, dp, dt);
    run<4>("inverse gs lazy4", d, dp, dt);
    run<5>("FAST exact mulhi + mad.lo nq", d, dp, dt);
    run<6>("FAST approx(asm) + C 64-bit mads", d, dp, dt);
    run<7>("FAST hand PTX (pack/unpack)", d, dp, dt);
    run<8>("guarded hand PTX", d, dp, dt);
    run<9>("FAST hand PTX, split IMAD chains", d, dp, dt);
    g_blocks_per_sm = 2;
    run<7>("FAST hand PTX @2 CTAs/SM (16 warps)", d, dp, dt);
    g_blocks_per_sm = 8;
    run<7>("FAST hand PTX @8 CTAs/SM (64 warps)", d, dp, dt);
    return 0;
}
