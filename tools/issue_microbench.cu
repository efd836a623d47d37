// tools/issue_microbench.cu -- issue-rate microbenchmarks for the instruction classes the modular butterflies use
// (B200 / sm_100a).  Every stream consists of 8 independent dependency chains per thread whose operands depend on the
// chain's previous result, so ptxas cannot hoist or strength-reduce anything; the SASS mix of every kernel is checked
// with `cuobjdump -sass` (see profiles/r02_issue_microbench.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_microbench issue_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITER 2048

#define WIDE(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
// operand taken from the chain itself (low word of the accumulator)
#define WIDE_SELF(i) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(y[i]));
#define WIDEC(i) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, hi, %1, %2;}" : "+l"(w[i]) : "r"(y[i]), "l"(w2[i]));
#define WIDEZ(i) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, hi, %1;}" : "+l"(w[i]) : "r"(y[i]));
#define IMADLO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
#define IMADHI(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
#define IADD3(i) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(z[i]) : "r"(y[i]), "r"(x[i]));
#define IADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(y[i]));
#define ADD64(i) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(z[i]), "+r"(v[i]) : "r"(y[i]), "r"(x[i]));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(y[i]), "r"(v[i]));
#define SHF(i) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[i]) : "r"(y[i]));
#define CSUB(i) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; @p sub.u32 %0, %0, %1;}" : "+r"(z[i]) : "r"(y[i]));
#define SELP(i) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(z[i]) : "r"(y[i]), "r"(v[i]));
#define DFMA(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fb[i]), "d"(fa[i]));
#define DADD(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(fb[i]));
#define DMUL(i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(fb[i]));
#define FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(gb[i]), "f"(ga[i]));
#define I2D(i) asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(f[i]) : "r"(z[i]));
#define D2I(i) asm volatile("cvt.rzi.u32.f64 %0, %1;" : "=r"(z[i]) : "d"(f[i]));
#define I2D64(i) asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(f[i]) : "l"(w[i]));
#define D2I64(i) asm volatile("cvt.rzi.u64.f64 %0, %1;" : "=l"(w[i]) : "d"(f[i]));
#define MULHI64(i) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w2[i]));
#define MULLO64(i) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w2[i]));

template <int KIND>
__global__ void __launch_bounds__(256) k(u32 *out, const u32 *in, long long *clk)
{
    long long t0 = clock64();
    u64 w[8], w2[8];
    u32 x[8], y[8], z[8], v[8];
    double f[8], fa[8], fb[8];
    float g[8], ga[8], gb[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        u32 s = in[threadIdx.x + 256 * i];
        x[i] = s * 3 + 1, y[i] = s ^ 0x9e3779b9u, z[i] = s + i, v[i] = s * 7 + i;
        w[i] = (u64)s * 0x9E3779B97F4A7C15ull, w2[i] = w[i] ^ 0x1234567ull;
        f[i] = s * 0.5 + i, fa[i] = s * 1e-9, fb[i] = 1.0 + s * 1e-12;
        g[i] = s * 0.25f + i, ga[i] = s * 1e-6f, gb[i] = 1.0f + s * 1e-9f;
    }
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (KIND == 0) { WIDE_SELF(i) }
            if (KIND == 1) { IMADLO(i) }
            if (KIND == 2) { IMADHI(i) }
            if (KIND == 3) { IADD3(i) }
            if (KIND == 4) { ADD64(i) }
            if (KIND == 5) { LOP(i) }
            if (KIND == 6) { SHF(i) }
            if (KIND == 7) { SELP(i) }
            if (KIND == 8) { WIDE_SELF(i) IADD3(i) }
            if (KIND == 9) { WIDE_SELF(i) ADD64(i) }
            if (KIND == 10) { WIDE_SELF(i) IADD3(i) LOP(i) }
            if (KIND == 11) { IMADLO(i) IADD3(i) }
            if (KIND == 12) { IMADLO(i) ADD64(i) }
            if (KIND == 13) { WIDE_SELF(i) IMADLO(i) }
            if (KIND == 14) { WIDE_SELF(i) IMADLO(i) ADD64(i) }
            if (KIND == 15) { WIDE_SELF(i) IMADLO(i) ADD64(i) SELP(i) }
            if (KIND == 16) { DFMA(i) }
            if (KIND == 17) { DADD(i) }
            if (KIND == 18) { DFMA(i) WIDE_SELF(i) }
            if (KIND == 19) { DFMA(i) IADD3(i) }
            if (KIND == 20) { DFMA(i) WIDE_SELF(i) IADD3(i) }
            if (KIND == 21) { FFMA(i) }
            if (KIND == 22) { FFMA(i) WIDE_SELF(i) }
            if (KIND == 23) { FFMA(i) IADD3(i) }
            if (KIND == 24) { I2D(i) D2I(i) }
            if (KIND == 25) { I2D64(i) D2I64(i) }
            if (KIND == 26) { MULHI64(i) }
            if (KIND == 27) { MULLO64(i) }
            if (KIND == 28) { WIDE(i) }
            if (KIND == 29) { WIDE_SELF(i) SELP(i) }
            if (KIND == 30) { WIDE_SELF(i) SHF(i) }
            if (KIND == 31) { WIDE_SELF(i) IADD3(i) IADD3(i) }
            if (KIND == 32) { DMUL(i) }
            if (KIND == 33) { IADD(i) }
            if (KIND == 34) { WIDEC(i) }
            if (KIND == 35) { WIDEZ(i) }
            if (KIND == 36) { WIDEC(i) IADD3(i) }
            if (KIND == 37) { WIDEZ(i) IADD3(i) }
            if (KIND == 38) { WIDEC(i) ADD64(i) }
            if (KIND == 39) { WIDEC(i) IMADLO(i) }
            if (KIND == 40) { WIDEC(i) IMADLO(i) ADD64(i) }
            if (KIND == 41) { WIDEC(i) IADD3(i) LOP(i) }
            if (KIND == 42) { WIDEC(i) SELP(i) }
            if (KIND == 43) { WIDEC(i) DFMA(i) }
            if (KIND == 44) { WIDEC(i) FFMA(i) }
            if (KIND == 45) { WIDEC(i) DFMA(i) IADD3(i) }
            if (KIND == 46) { IMADLO(i) DFMA(i) }
            if (KIND == 47) { WIDEC(i) IADD3(i) IADD3(i) }
        }
        if (KIND == 28)
        {
            // keep the WIDE operands chain-dependent without adding instructions to the measured count: rotate roles
#pragma unroll
            for (int i = 0; i < 8; i++)
                x[i] = (u32)w[(i + 1) & 7];
        }
    }
    // only the arrays a stream touches stay live across the loop (register pressure decides the resident warps)
    constexpr bool USES_W = KIND == 0 || (KIND >= 8 && KIND <= 10) || (KIND >= 13 && KIND <= 15) || KIND == 18 || KIND == 20 || KIND == 22 ||
                            (KIND >= 25 && KIND <= 31) || (KIND >= 34 && KIND <= 45) || KIND == 47;
    constexpr bool USES_F = (KIND >= 16 && KIND <= 20) || KIND == 24 || KIND == 25 || KIND == 32 || KIND == 43 || KIND == 45 || KIND == 46;
    constexpr bool USES_G = (KIND >= 21 && KIND <= 23) || KIND == 44;
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        s += x[i] + y[i] + z[i] + v[i];
        if (USES_W)
            s += (u32)w[i] + (u32)(w[i] >> 32) + (u32)w2[i];
        if (USES_F)
            s += (u32)__double2uint_rz(f[i]);
        if (USES_G)
            s += (u32)__float2uint_rz(g[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0)
        clk[blockIdx.x] = clock64() - t0;
}

static int g_ctas_per_sm = 8;
template <int KIND>
void run(const char *name, int ops_per_iter)
{
    u32 *d, *in;
    const int blocks = 148 * g_ctas_per_sm;
    cudaMalloc(&d, blocks * 256 * 4);
    cudaMalloc(&in, 2048 * 4);
    cudaMemset(in, 0x5a, 2048 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    long long *dclk;
    cudaMalloc(&dclk, blocks * 8);
    k<KIND><<<blocks, 256>>>(d, in, dclk);
    cudaEventRecord(e0);
    k<KIND><<<blocks, 256>>>(d, in, dclk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    static long long hclk[148 * 16];
    cudaMemcpy(hclk, dclk, blocks * 8, cudaMemcpyDeviceToHost);
    double cycles = 0; // SM clocks the resident CTAs ran for (all CTAs of the grid are co-resident)
    for (int i = 0; i < blocks; i++)
        cycles += (double)hclk[i] / blocks;
    cudaFree(dclk);
    double warp_instr = (double)blocks * 8 /*warps*/ * ITER * 8.0 * ops_per_iter;
    // event time x SM clock = cycles the whole grid took; per SMSP one "group" = one warp executing the stream once
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double groups_per_smsp = (double)blocks * 8 * ITER * 8.0 / (148.0 * 4);
    int regs = 0;
    {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k<KIND>);
        regs = fa.numRegs;
    }
    (void)cycles, (void)warp_instr;
    printf("%-40s %8.3f ms  %6.2f clk per group per SMSP (nominal %d instr)  regs %d\n", name, ms, ms * 1e-3 * khz * 1e3 / groups_per_smsp, ops_per_iter, regs);
    cudaFree(d);
    cudaFree(in);
}
int main(int argc, char **argv)
{
    if (argc > 1)
        g_ctas_per_sm = atoi(argv[1]);
    printf("CTAs/SM = %d (256 threads each)\n", g_ctas_per_sm);
    run<0>("IMAD.WIDE (chain operand)", 1);
    run<28>("IMAD.WIDE (register operands)", 1);
    run<34>("IMAD.WIDE Rc addend (pure)", 1);
    run<35>("IMAD.WIDE RZ addend (pure)", 1);
    run<36>("WIDEC + IADD3", 2);
    run<47>("WIDEC + 2 IADD3", 3);
    run<37>("WIDEZ + IADD3", 2);
    run<38>("WIDEC + add64", 3);
    run<39>("WIDEC + IMAD", 2);
    run<40>("WIDEC + IMAD + add64", 4);
    run<41>("WIDEC + IADD3 + LOP3", 3);
    run<42>("WIDEC + ISETP/SEL", 3);
    run<43>("WIDEC + DFMA", 2);
    run<44>("WIDEC + FFMA", 2);
    run<45>("WIDEC + DFMA + IADD3", 3);
    run<46>("IMAD + DFMA", 2);
    run<1>("IMAD lo", 1);
    run<2>("IMAD.HI", 1);
    run<33>("IADD (2-input)", 1);
    run<3>("IADD3 (2 adds -> 1 IADD3?)", 1);
    run<4>("64-bit add (IADD3 + IADD3.X)", 2);
    run<5>("LOP3", 1);
    run<6>("SHF", 1);
    run<7>("ISETP + SEL", 2);
    run<8>("WIDE + IADD3", 2);
    run<31>("WIDE + 2 IADD3", 3);
    run<9>("WIDE + add64", 3);
    run<10>("WIDE + IADD3 + LOP3", 3);
    run<11>("IMAD + IADD3", 2);
    run<12>("IMAD + add64", 3);
    run<13>("WIDE + IMAD", 2);
    run<14>("WIDE + IMAD + add64", 4);
    run<15>("WIDE + IMAD + add64 + ISETP/SEL", 6);
    run<29>("WIDE + ISETP/SEL", 3);
    run<30>("WIDE + SHF", 2);
    run<16>("DFMA", 1);
    run<17>("DADD", 1);
    run<32>("DMUL", 1);
    run<18>("DFMA + WIDE", 2);
    run<19>("DFMA + IADD3", 2);
    run<20>("DFMA + WIDE + IADD3", 3);
    run<21>("FFMA", 1);
    run<22>("FFMA + WIDE", 2);
    run<23>("FFMA + IADD3", 2);
    run<24>("I2F.F64.U32 + F2I.U32.F64", 2);
    run<25>("I2F.F64.U64 + F2I.U64.F64", 2);
    run<26>("mul.hi.u64", 1);
    run<27>("mul.lo.u64", 1);
    return 0;
}
