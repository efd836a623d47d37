// seal_b200/csrc/sb_src.cuh -- how the key-switching kernels see their operands (shared by sb_engine.cu and sb_ksint.cu).
#pragma once
#include "sb_device.cuh"

namespace sb
{
    // --------------------------------------------------------------------------------- source accessors ----
    // A batch of [L][n] polynomials, optionally seen through a Galois automorphism (galois.cpp:148-218).
    struct Src
    {
        const u64 *p = nullptr;
        long long bstride = 0; // words between consecutive batch items
        const uint32_t *perm = nullptr; // NTT-form: out[i] = in[perm[i]]
        uint32_t ginv = 0;              // coefficient form: inverse Galois element mod 2n (0 = identity)
        int logn = 0;

        __device__ __forceinline__ bool plain() const { return perm == nullptr && ginv == 0; }
        __device__ __forceinline__ const u64 *row(int b, int J) const { return p + b * bstride + (static_cast<long long>(J) << logn); }
        __device__ __forceinline__ u64 get(int b, int J, int idx, u64 qJ) const
        {
            const u64 *r = p + b * bstride + (static_cast<long long>(J) << logn);
            if (perm)
                return r[perm[idx]];
            if (ginv)
            {
                uint32_t ip = (static_cast<uint32_t>(idx) * ginv) & ((2u << logn) - 1u);
                u64 v = r[ip & ((1u << logn) - 1u)];
                if (ip >> logn)
                    v = v ? qJ - v : 0;
                return v;
            }
            return r[idx];
        }
    };


    // what the mod-down result is added to (the ciphertext being updated)
    struct BaseSrc
    {
        Src s;            // polys 0,1 of the base ciphertext: poly c of item b at s.p + b*s.bstride + c*pstride
        long long pstride = 0;
        int c1_zero = 0;  // apply_galois: component 1 starts from zero (evaluator.cpp:2487)
        int present = 0;
        __device__ __forceinline__ u64 get(int b, int c, int i, int idx, u64 q) const
        {
            if (!present || (c == 1 && c1_zero))
                return 0;
            Src t = s;
            t.p += c * pstride;
            return t.get(b, i, idx, q);
        }
    };

} // namespace sb
