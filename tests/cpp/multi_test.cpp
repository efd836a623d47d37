// tests/cpp/multi_test.cpp -- the C++ multi-device dispatcher (sb200_group_*, seal_b200/csrc/sb_multi.cpp) against the
// single-device entry points, through the C-ABI only (no reference needed): contiguous slices over the devices of the box give
// the same words as one device doing the whole batch.  With one visible GPU the group is built from two contexts on that GPU,
// which exercises the same slicing and threading.
#include "seal_b200.h"
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                     \
    do                                                                  \
    {                                                                   \
        g_checks++;                                                     \
        if (!(cond))                                                    \
        {                                                               \
            g_fail++;                                                   \
            std::printf("FAIL %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, sb200_last_error()); \
        }                                                               \
    } while (0)

int main(int argc, char **argv)
{
    const size_t n = 8192, k = 4, L = 3, batch = 7;
    int bits[k] = { 54, 54, 54, 55 };
    uint64_t q[k];
    CHECK(sb200_coeff_modulus_create(n, bits, k, q) == SB200_OK);
    int ndev = argc > 1 ? std::atoi(argv[1]) : 0; // 0: every visible device
    std::vector<int> devs;
    sb200_group *g = nullptr;
    if (ndev == 1)
        devs = { 0, 0 }; // two contexts on one GPU
    else
        for (int d = 0; d < ndev; d++)
            devs.push_back(d);
    CHECK(sb200_group_create(SB200_SCHEME_CKKS, n, q, k, 0, devs.empty() ? nullptr : devs.data(), devs.size(), &g) == SB200_OK);
    if (!g)
        return 2;
    const size_t G = sb200_group_size(g);
    std::printf("group of %zu contexts\n", G);
    size_t covered = 0;
    for (size_t i = 0; i < G; i++)
    {
        size_t f, c;
        CHECK(sb200_group_slice(g, batch, i, &f, &c) == SB200_OK);
        CHECK(f == covered);
        covered += c;
    }
    CHECK(covered == batch);

    std::mt19937_64 rng(0x5EA1);
    auto fill = [&](std::vector<uint64_t> &v, size_t rows_per_item, size_t items, size_t nprimes) {
        v.resize(items * rows_per_item * nprimes * n);
        for (size_t it = 0; it < items * rows_per_item; it++)
            for (size_t i = 0; i < nprimes; i++)
                for (size_t j = 0; j < n; j++)
                    v[(it * nprimes + i) * n + j] = rng() % q[i];
    };
    std::vector<uint64_t> key, a, b;
    fill(key, 2, L, k);
    fill(a, 2, batch, L);
    fill(b, 2, batch, L);
    sb200_group_key *gk = nullptr;
    CHECK(sb200_group_kswitch_key_create(g, key.data(), L, &gk) == SB200_OK);
    sb200_context *c0 = sb200_group_context(g, 0);
    sb200_kswitch_key *k0 = nullptr;
    CHECK(sb200_kswitch_key_create(c0, key.data(), L, &k0) == SB200_OK);

    const size_t w = 2 * L * n;
    std::vector<uint64_t> one(batch * w), many(batch * w, 1);
    CHECK(sb200_multiply_relinearize_host(c0, L, batch, a.data(), b.data(), k0, one.data()) == SB200_OK);
    CHECK(sb200_group_multiply_relinearize_host(g, L, batch, a.data(), b.data(), gk, many.data()) == SB200_OK);
    CHECK(one == many);
    std::vector<uint64_t> r1(batch * 2 * (L - 1) * n), r2(r1.size(), 1);
    CHECK(sb200_rescale_to_next_host(c0, L, batch, one.data(), r1.data()) == SB200_OK);
    CHECK(sb200_group_rescale_to_next_host(g, L, batch, one.data(), r2.data()) == SB200_OK);
    CHECK(r1 == r2);
    std::vector<uint64_t> m1(batch * 3 * L * n), m2(m1.size(), 1), l1(batch * w), l2(batch * w, 1);
    CHECK(sb200_multiply_host(c0, L, batch, a.data(), b.data(), m1.data()) == SB200_OK);
    CHECK(sb200_group_multiply_host(g, L, batch, a.data(), b.data(), m2.data()) == SB200_OK);
    CHECK(m1 == m2);
    CHECK(sb200_group_relinearize_host(g, L, batch, m2.data(), gk, l2.data()) == SB200_OK);
    CHECK(l2 == one); // relinearize(multiply) == fused
    std::vector<uint64_t> t1 = a, t2 = a;
    CHECK(sb200_ntt_inverse_host(c0, L, 2, batch, t1.data()) == SB200_OK);
    CHECK(sb200_group_ntt_inverse_host(g, L, 2, batch, t2.data()) == SB200_OK);
    CHECK(t1 == t2);
    CHECK(sb200_group_ntt_forward_host(g, L, 2, batch, t2.data()) == SB200_OK);
    CHECK(t2 == a);
    // a batch smaller than the group: the empty slices are skipped
    std::vector<uint64_t> s1(w), s2(w, 1);
    CHECK(sb200_multiply_relinearize_host(c0, L, 1, a.data(), b.data(), k0, s1.data()) == SB200_OK);
    CHECK(sb200_group_multiply_relinearize_host(g, L, 1, a.data(), b.data(), gk, s2.data()) == SB200_OK);
    CHECK(s1 == s2);
    // errors surface with the single-device status codes
    CHECK(sb200_group_multiply_relinearize_host(g, L, 0, a.data(), b.data(), gk, s2.data()) == SB200_E_INVALID_ARG);
    CHECK(sb200_group_rescale_to_next_host(g, 1, batch, one.data(), r2.data()) == SB200_E_INVALID_ARG);
    sb200_kswitch_key_destroy(k0);
    sb200_group_kswitch_key_destroy(gk);
    sb200_group_destroy(g);
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAIL" : "PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
