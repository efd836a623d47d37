// seal_b200/csrc/sb_device.cuh -- device-side modular arithmetic and shared structs (sm_100a).
//
// All residues are uint64 (primes < 2^61: user primes <= 60 bits, BEHZ auxiliary primes 61 bits; reference limits in
// util/defines.h:32-71).  Twiddle multiplications use Shoup/Harvey precomputed quotients (the same mathematical device
// the reference uses in util/uintarithsmallmod.h:255-326), general products use a 128-bit Barrett step.  Outputs that
// leave a kernel are canonical residues in [0, q) so results are word-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

struct __align__(16) Tw
{
    u64 w;  // twiddle value in [0, q)
    u64 wq; // floor(w * 2^64 / q)
};

// Per-prime constants + table pointers, one entry per prime id (q-chain primes first, then BEHZ auxiliary primes).
struct PrimeDev
{
    u64 q;
    u64 q2;         // 2q
    u64 ratio_lo;   // floor(2^128 / q), low word
    u64 ratio_hi;   //                   high word (= floor(2^64 / q))
    Tw inv_n;       // n^-1 mod q
    Tw inv_n_w;     // n^-1 * (last inverse-stage twiddle) mod q
    const Tw *fwd;  // fwd[m + i]: forward stage with m groups, group i  (psi powers, bit-reversed order)
    const Tw *inv;  // inv[m + i]: inverse stage with m groups, group i  (psi^-1 powers)
};

__device__ __forceinline__ u64 csub(u64 x, u64 q)
{
    return x >= q ? x - q : x;
}

// x * w mod q, lazily: result in [0, 2q) for ANY 64-bit x (w < q, wq = floor(w 2^64 / q)).
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, Tw t, u64 q)
{
    u64 h = __umul64hi(x, t.wq);
    return x * t.w - h * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, Tw t, u64 q)
{
    return csub(mul_shoup_lazy(x, t, q), q);
}

// x mod q for any 64-bit x (ratio_hi = floor(2^64 / q)); result canonical.
__device__ __forceinline__ u64 barrett64(u64 x, u64 q, u64 ratio_hi)
{
    u64 t = __umul64hi(x, ratio_hi);
    return csub(x - t * q, q);
}

// (hi:lo) mod q for a 128-bit value; ratio = floor(2^128 / q).  Exact floor of the 256-bit product's top limb is
// approximated from below by at most 1, so one conditional subtraction canonicalises.
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 ratio_lo, u64 ratio_hi)
{
    u64 c0 = __umul64hi(lo, ratio_lo);
    u64 a_lo = lo * ratio_hi, a_hi = __umul64hi(lo, ratio_hi);
    u64 b_lo = hi * ratio_lo, b_hi = __umul64hi(hi, ratio_lo);
    u64 s = a_lo + c0;
    u64 carry = (s < a_lo);
    u64 s2 = s + b_lo;
    carry += (s2 < s);
    u64 t = hi * ratio_hi + a_hi + b_hi + carry;
    return csub(lo - t * q, q);
}

// a * b mod q, canonical, a and b arbitrary 64-bit with a*b < 2^128 trivially.
__device__ __forceinline__ u64 mulmod_barrett(u64 a, u64 b, const PrimeDev &P)
{
    return barrett128(a * b, __umul64hi(a, b), P.q, P.ratio_lo, P.ratio_hi);
}

// 128-bit accumulate: (hi:lo) += a*b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b)
{
    u64 pl = a * b, ph = __umul64hi(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

// Harvey butterflies on lazily reduced values.
// forward (Cooley-Tukey): inputs in [0,4q) -> outputs in [0,4q)
__device__ __forceinline__ void ct_bfly(u64 &x, u64 &y, Tw w, u64 q, u64 q2)
{
    u64 u = csub(x, q2);
    u64 v = mul_shoup_lazy(y, w, q);
    x = u + v;
    y = u - v + q2;
}
// inverse (Gentleman-Sande): inputs in [0,2q) -> outputs in [0,2q)
__device__ __forceinline__ void gs_bfly(u64 &x, u64 &y, Tw w, u64 q, u64 q2)
{
    u64 u = x, v = y;
    x = csub(u + v, q2);
    y = mul_shoup_lazy(u - v + q2, w, q);
}

__device__ __forceinline__ Tw ldg_tw(const Tw *p)
{
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    Tw t;
    t.w = v.x;
    t.wq = v.y;
    return t;
}

// ---- TMA (1-D bulk async copy global -> shared) + mbarrier helpers --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(u64 *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, uint32_t phase)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(phase)
                 : "memory");
}
