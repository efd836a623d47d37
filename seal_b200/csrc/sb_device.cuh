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
    u64 q4;         // 4q
    u64 nq;         // 2^64 - q
    u64 zero;       // always 0, but not a compile-time constant: see ct_bfly
    u64 ratio_lo;   // floor(2^128 / q), low word
    u64 ratio_hi;   //                   high word (= floor(2^64 / q))
    Tw inv_n;       // n^-1 mod q
    Tw inv_n_w;     // n^-1 * (last inverse-stage twiddle) mod q
    const Tw *fwd;  // fwd[m + i]: forward stage with m groups, group i  (psi powers, bit-reversed order)
    const Tw *inv;  // inv[m + i]: inverse stage with m groups, group i  (psi^-1 powers)
    // q = 2^bits - dsol: every prime CoeffModulus::Create / get_primes makes has this shape with a small dsol
    // (util/numth.cpp:278-311: 2^bits - c * 2n + 1); dsol == 0 marks a prime without it (no Solinas folding)
    unsigned bits, dsol;
};

__device__ __forceinline__ u64 csub(u64 x, u64 q)
{
    return x >= q ? x - q : x;
}

// ---- Shoup / Harvey multiplication tuned for the sm_100 integer pipes -------------------------------------------
// Integer multiplies issue on the FMA-heavy pipe only (ncu: the binding pipe of every transform kernel) and a
// 32x32->64 multiply-add costs twice a 32-bit one, so the quotient estimate uses three wide products instead of the
// four (+ carry chain) of an exact 64x64 high half:
//     T = y1*wq1 + hi32(y1*wq0) + hi32(y0*wq1)      with  t-2 <= T <= t,  t = floor(y*wq / 2^64)
// and the remainder y*w - T*q is formed as lo64(y*w + T*(2^64-q)) with two wide and four 32-bit multiply-adds.
// Result: x*w mod q in [0, 4q) for ANY 64-bit x (w < q, wq = floor(w 2^64 / q), q < 2^61).
__device__ __forceinline__ u64 pack64(unsigned lo, unsigned hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack64(u64 v, unsigned &lo, unsigned &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
// written so that ptxas keeps the sum of the two 32-bit high halves in ONE three-input 64-bit add (IADD3 with two carry outputs +
// IADD3.X) on the integer ALU instead of materialising the carry (SEL) and adding it on the multiply pipe (IMAD.X)
__device__ __forceinline__ u64 approx_mulhi(u64 y, u64 wq)
{
    unsigned y0, y1, wq0, wq1, alo, ahi, blo, bhi;
    unpack64(y, y0, y1);
    unpack64(wq, wq0, wq1);
    u64 a, b, T;
    // the two cross products have equal weight 2^32: keep their high halves, drop their low halves and y0*wq0
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(y1), "r"(wq0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(y0), "r"(wq1));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(T) : "r"(y1), "r"(wq1));
    unpack64(a, alo, ahi);
    unpack64(b, blo, bhi);
    return T + static_cast<u64>(ahi) + static_cast<u64>(bhi);
}
// lo64(y*w + T*nq)
__device__ __forceinline__ u64 mullo_combine(u64 y, u64 w, u64 T, u64 nq)
{
    unsigned y0, y1, w0, w1, T0, T1, n0, n1, lo, hi;
    unpack64(y, y0, y1);
    unpack64(w, w0, w1);
    unpack64(T, T0, T1);
    unpack64(nq, n0, n1);
    u64 acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(T0), "r"(n0));
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(T1), "r"(n0));
    return pack64(lo, hi);
}
__device__ __forceinline__ u64 mul_shoup_lazy4(u64 x, Tw t, u64 nq)
{
    return mullo_combine(x, t.w, approx_mulhi(x, t.wq), nq);
}
// exact variants (result in [0,2q) / canonical) for the element-wise epilogues
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, Tw t, u64 q)
{
    u64 h = __umul64hi(x, t.wq);
    return x * t.w - h * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, Tw t, u64 q)
{
    return csub(mul_shoup_lazy(x, t, q), q);
}
// x mod q, lazily in [0, 4q), for any 64-bit x (three wide products + one wide + two 32-bit multiply-adds)
__device__ __forceinline__ u64 barrett_lazy4(u64 x, u64 ratio_hi, u64 nq)
{
    u64 T = approx_mulhi(x, ratio_hi);
    const unsigned T0 = static_cast<unsigned>(T), T1 = static_cast<unsigned>(T >> 32);
    const unsigned n0 = static_cast<unsigned>(nq), n1 = static_cast<unsigned>(nq >> 32);
    u64 acc = static_cast<u64>(T0) * n0 + x;
    unsigned hi = static_cast<unsigned>(acc >> 32) + T0 * n1 + T1 * n0;
    return (static_cast<u64>(hi) << 32) | static_cast<unsigned>(acc);
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
    // one carry chain instead of a compare + select per accumulate
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\tmadc.hi.u64 %1, %2, %3, %1;" : "+l"(lo), "+l"(hi) : "l"(a), "l"(b));
}

// ---- key multiply-accumulate on 28-bit limbs (fused key-switch kernel) ------------------------------------------------
// A carry-chained 128-bit multiply-accumulate costs ~7 wide multiplies + ~10 adds as ptxas expands it.  For primes
// q = 2^bits - dsol below 2^56 the transformed digit is first folded below 2^56 (Solinas: x = (x mod 2^bits) + (x >> bits) * dsol,
// one wide multiply), split into 28-bit limbs a = a1*2^28 + a0, and multiplied with the key word kept in the same limb form
// k = k1*2^28 + k0 (encoded once at key upload) Karatsuba-style into three plain 64-bit column sums without any carry:
//     S0 += a0*k0      S2 += a1*k1      S1 += (a0+a1)*(k0+k1)          (each product < 2^58; up to 60 digits per sum)
// and the 128-bit value S0 + (S1 - S0 - S2)*2^28 + S2*2^56 is formed once at the end: 3 wide multiplies per product.
__device__ __forceinline__ u64 fold_solinas(u64 a, unsigned bits, unsigned dsol)
{
    const u64 low = a & ((1ull << bits) - 1ull);
    return low + static_cast<u64>(static_cast<unsigned>(a >> bits)) * dsol;
}
__device__ __forceinline__ u64 limb28_encode(u64 k)
{
    return (k & 0x0FFFFFFFull) | ((k >> 28) << 32);
}
__device__ __forceinline__ u64 limb28_decode(u64 e)
{
    return (e & 0xFFFFFFFFull) | ((e >> 32) << 28);
}
struct Acc3
{
    u64 s0, s1, s2;
};
// a0, a1, as = a0 + a1 of the folded digit; ke = limb-encoded key word
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
    const u64 mid = s.s1 - s.s0 - s.s2; // sum of a0*k1 + a1*k0, < 2^63
    // s0 + mid * 2^28 + s2 * 2^56
    u64 l = s.s0, h = 0, t;
    t = mid << 28;
    l += t, h += (mid >> 36) + (l < t);
    t = s.s2 << 56;
    l += t, h += (s.s2 >> 8) + (l < t);
    lo = l, hi = h;
}

// Harvey-style butterflies on lazily reduced values.
// forward (Cooley-Tukey).  FAST (all primes of the launch < 2^57): no conditional subtraction at all -- values grow
//   by at most 4q per stage ((4 + 4*17) q < 2^64), the kernel reduces once before the store.
//   guarded (any prime < 2^61): inputs in [0,8q) -> outputs in [0,8q).
template <bool FAST>
__device__ __forceinline__ void ct_bfly(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 v = mul_shoup_lazy4(y, w, P.nq);
    u64 u = FAST ? x : csub(x, P.q4);
    // three-operand form: keeps the 64-bit adds on the integer ALU (IADD3/IADD3.X); the two-operand form is scheduled by
    // ptxas onto the multiply pipe (IMAD.X), which is the pipe that limits these kernels
    x = u + v + P.zero;
    y = u - v + P.q4;
}
// inverse (Gentleman-Sande): inputs in [0,4q) -> outputs in [0,4q)
__device__ __forceinline__ void gs_bfly(u64 &x, u64 &y, Tw w, const PrimeDev &P)
{
    u64 u = x, v = y;
    x = csub(u + v, P.q4);
    y = mul_shoup_lazy4(u - v + P.q4, w, P.nq);
}
// forward values before the store: -> [0, 4q)
template <bool FAST>
__device__ __forceinline__ u64 fwd_finish(u64 v, const PrimeDev &P)
{
    return FAST ? barrett_lazy4(v, P.ratio_hi, P.nq) : csub(v, P.q4);
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
