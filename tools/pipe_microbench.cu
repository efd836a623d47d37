// tools/pipe_microbench.cu -- measures issue throughput of the integer instructions the NTT butterflies are made of
// (B200 / sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITER 4096
template <int KIND>
__global__ void k(u32 *out, u32 a0, u32 b0)
{
    u32 a = a0 + threadIdx.x, b = b0 | 1;
    u64 w[8];
    u32 r[8], r2[8], r3[8];
    u64 w2[8];
    double f[8], fa = a0 * 1.0000001, fb = 1.0 + b0 * 1e-9;
    float g[8], ga = a0 * 1.5f, gb = 1.0f + b0 * 1e-6f;
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a * (i + 3), r[i] = a ^ (i * 77), r2[i] = a + i, r3[i] = a - i, w2[i] = w[i] * 0x9E3779B97F4A7C15ull + i, f[i] = a * 0.5 + i, g[i] = a * 0.25f + i;
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a));
            if (KIND == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(r[i]), "r"(b));
            if (KIND == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a));
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
            if (KIND == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(a));
            if (KIND == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(((u32*)w)[2*i]) : "r"(b)); }
            if (KIND == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); }
            if (KIND == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i+1)&7]) : "r"(b), "r"(a)); }
            if (KIND == 8) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); asm volatile("addc.u32 %0, %0, %1;" : "+r"(((u32*)w)[2*i]) : "r"(a)); }
            if (KIND == 9) { asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(r[i]) : "r"(b), "r"(a)); }
            if (KIND == 10) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); }
            if (KIND == 11) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(r[i]), "r"(b)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(((u32*)w)[2*i+1])); }
            if (KIND == 12) { asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); }

            if (KIND == 14) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fb), "d"(fa));
            if (KIND == 15) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fb), "d"(fa)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); }
            if (KIND == 16) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(gb), "f"(ga));
            if (KIND == 17) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(gb), "f"(ga)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); }
            if (KIND == 18) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(r[i]), "+r"(r2[i]) : "r"(b), "r"(a)); }
            if (KIND == 19) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(b)); }
            if (KIND == 20) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(b)); }
            if (KIND == 21) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2[i]) : "r"(b), "r"(a)); asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(r[i]), "+r"(r2[i]) : "r"(b), "r"(a)); }
            if (KIND == 22) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(r2[i]), "+r"(r3[i]) : "r"(b), "r"(a)); }
            if (KIND == 23) { asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w2[(i+1)&7])); }
            if (KIND == 24) { asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w2[(i+1)&7])); }
            if (KIND == 25) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(r[i]), "r"(r2[i])); }
            if (KIND == 26) { asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r2[i]) : "r"(b), "r"(a)); }
            if (KIND == 27) { asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); }
            if (KIND == 28) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(f[i]) : "r"(r[i])); asm volatile("cvt.rzi.u32.f64 %0, %1;" : "=r"(r[i]) : "d"(f[i])); }
            if (KIND == 29) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(fb)); }
            if (KIND == 30) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(fb), "d"(fa)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); }
            if (KIND == 13) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[(i+1)&7]) : "r"(b), "r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[(i+3)&7]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(((u32*)w)[2*((i+2)&7)]) : "r"(b)); }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i] + r2[i] + r3[i] + (u32)w[i] + (u32)(w[i] >> 32) + (u32)w2[i] + (u32)(w2[i] >> 32) + (u32)__double2uint_rz(f[i]) + (u32)__float2uint_rz(g[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run(const char *name, int ops_per_iter)
{
    u32 *d;
    cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(d, 5, 7);
    cudaEventRecord(e0);
    k<KIND><<<148 * 8, 256>>>(d, 5, 7);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warp_instr = 148.0 * 8 * 8 /*warps*/ * ITER * 8.0 * ops_per_iter;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM  (%.2f per SMSP)\n", name, ms, warp_instr / cycles / 148.0, warp_instr / cycles / 148.0 / 4);
    cudaFree(d);
}
int main()
{
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.WIDE.U32", 1);
    run<2>("IMAD.HI.U32", 1);
    run<3>("IADD", 1);
    run<4>("LOP3", 1);
    run<5>("IMAD + IADD", 2);
    run<6>("IMAD.WIDE + IADD", 2);
    run<7>("IMAD.WIDE + IADD + LOP3", 3);
    run<8>("IADD.CC + IADDC", 2);
    run<9>("ISETP + SEL", 2);
    run<10>("IMAD.WIDE + IMAD", 2);
    run<11>("MUL.WIDE (no addend) + XOR", 2);
    run<12>("IMUL lo (no addend)", 1);
    run<13>("2 WIDE + 2 IMAD + 1 IADD", 5);
    run<14>("DFMA", 1);
    run<15>("DFMA + IMAD.WIDE", 2);
    run<16>("FFMA", 1);
    run<17>("FFMA + IMAD.WIDE", 2);
    run<18>("IMAD.WIDE + IADD.CC + IADDC", 3);
    run<19>("IMAD.WIDE + SHF", 2);
    run<20>("SHF", 1);
    run<21>("2 WIDE + IADD.CC + IADDC", 4);
    run<22>("IMAD + IADD.CC + IADDC", 3);
    run<23>("mul.hi.u64 (expansion)", 1);
    run<24>("mul.lo.u64 (expansion)", 1);
    run<25>("IMAD.WIDE all-reg operands", 1);
    run<26>("IADD + LOP3", 2);
    run<27>("ISETP + SEL + IMAD.WIDE", 3);
    run<28>("I2F.F64.U32 + F2I.U32.F64", 2);
    run<29>("DADD", 1);
    run<30>("DFMA + IADD", 2);
    return 0;
}
