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
    u32 r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a * (i + 3), r[i] = a ^ (i * 77);
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
            if (KIND == 13) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[(i+1)&7]) : "r"(b), "r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[(i+3)&7]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(((u32*)w)[2*((i+2)&7)]) : "r"(b)); }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i] + (u32)w[i] + (u32)(w[i] >> 32);
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
    return 0;
}
