// seal_b200/csrc/sb_multi.cpp -- multi-device dispatch on top of the single-device C-ABI (include/seal_b200.h, "sb200_group_*").
// Host C++ only: one context per device, one host thread per device for the duration of a call, contiguous slices of the batch,
// no data-path collective (every ciphertext of a batch is independent; tables and keys are replicated).
#include "../../include/seal_b200.h"
#include <cuda_runtime_api.h>
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct sb200_group
{
    std::vector<sb200_context *> ctx;
    size_t n = 0;
    ~sb200_group()
    {
        for (auto c : ctx)
            if (c)
                sb200_context_destroy(c);
    }
};
struct sb200_group_key
{
    std::vector<sb200_kswitch_key *> key;
    ~sb200_group_key()
    {
        for (auto k : key)
            if (k)
                sb200_kswitch_key_destroy(k);
    }
};

namespace
{
    // contiguous, balanced: the first (batch % g) slots get one ciphertext more
    void slice(size_t batch, size_t g, size_t i, size_t &first, size_t &count)
    {
        const size_t base = batch / g, extra = batch % g;
        count = base + (i < extra ? 1 : 0);
        first = i * base + (i < extra ? i : extra);
    }
    // runs fn(slot, first, count) on one thread per device; the first failing status wins.  The error TEXT lives in the failing
    // thread's thread-local slot of the single-device layer, so it is re-raised here by repeating the failing call's status only.
    template <class F>
    int fan_out(const sb200_group *g, size_t batch, F &&fn)
    {
        if (!g)
            return SB200_E_POINTER;
        if (batch == 0)
            return SB200_E_INVALID_ARG;
        const size_t G = g->ctx.size();
        std::vector<int> rc(G, SB200_OK);
        std::vector<std::thread> th;
        for (size_t i = 0; i < G; i++)
        {
            size_t first, count;
            slice(batch, G, i, first, count);
            if (!count)
                continue;
            th.emplace_back([&, i, first, count] { rc[i] = fn(i, first, count); });
        }
        for (auto &t : th)
            t.join();
        for (int r : rc)
            if (r != SB200_OK)
                return r;
        return SB200_OK;
    }
} // namespace

extern "C" {

int sb200_group_create(int scheme, size_t n, const uint64_t *coeff_modulus, size_t k, uint64_t plain_modulus, const int *devices, size_t device_count,
                       sb200_group **out)
{
    if (!coeff_modulus || !out)
        return SB200_E_POINTER;
    std::vector<int> devs;
    if (devices)
        devs.assign(devices, devices + device_count);
    else
    {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0)
            return SB200_E_CUDA;
        for (int d = 0; d < cnt; d++)
            devs.push_back(d);
    }
    if (devs.empty())
        return SB200_E_INVALID_ARG;
    auto g = std::make_unique<sb200_group>();
    g->n = n;
    g->ctx.assign(devs.size(), nullptr);
    // contexts are built concurrently: table generation is host work (seconds at n = 65536)
    std::vector<int> rc(devs.size(), SB200_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < devs.size(); i++)
        th.emplace_back([&, i] { rc[i] = sb200_context_create(scheme, n, coeff_modulus, k, plain_modulus, devs[i], &g->ctx[i]); });
    for (auto &t : th)
        t.join();
    for (size_t i = 0; i < devs.size(); i++)
        if (rc[i] != SB200_OK)
        {
            // repeat the failing creation on this thread so that sb200_last_error() carries its message here
            sb200_context *tmp = nullptr;
            const int r = sb200_context_create(scheme, n, coeff_modulus, k, plain_modulus, devs[i], &tmp);
            if (tmp)
                sb200_context_destroy(tmp);
            return r != SB200_OK ? r : rc[i];
        }
    *out = g.release();
    return SB200_OK;
}

int sb200_group_destroy(sb200_group *group)
{
    if (!group)
        return SB200_E_POINTER;
    delete group;
    return SB200_OK;
}

size_t sb200_group_size(const sb200_group *group)
{
    return group ? group->ctx.size() : 0;
}

sb200_context *sb200_group_context(sb200_group *group, size_t i)
{
    return group && i < group->ctx.size() ? group->ctx[i] : nullptr;
}

int sb200_group_slice(const sb200_group *group, size_t batch, size_t i, size_t *first, size_t *count)
{
    if (!group || !first || !count)
        return SB200_E_POINTER;
    if (i >= group->ctx.size())
        return SB200_E_OUT_OF_RANGE;
    slice(batch, group->ctx.size(), i, *first, *count);
    return SB200_OK;
}

int sb200_group_kswitch_key_create(sb200_group *group, const uint64_t *h_key, size_t digits, sb200_group_key **out)
{
    if (!group || !h_key || !out)
        return SB200_E_POINTER;
    auto k = std::make_unique<sb200_group_key>();
    k->key.assign(group->ctx.size(), nullptr);
    std::vector<int> rc(group->ctx.size(), SB200_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < group->ctx.size(); i++)
        th.emplace_back([&, i] { rc[i] = sb200_kswitch_key_create(group->ctx[i], h_key, digits, &k->key[i]); });
    for (auto &t : th)
        t.join();
    for (size_t i = 0; i < rc.size(); i++)
        if (rc[i] != SB200_OK)
        {
            sb200_kswitch_key *tmp = nullptr;
            const int r = sb200_kswitch_key_create(group->ctx[i], h_key, digits, &tmp); // for the message
            if (tmp)
                sb200_kswitch_key_destroy(tmp);
            return r != SB200_OK ? r : rc[i];
        }
    *out = k.release();
    return SB200_OK;
}

int sb200_group_kswitch_key_destroy(sb200_group_key *key)
{
    if (!key)
        return SB200_E_POINTER;
    delete key;
    return SB200_OK;
}

#define NEED3(a, b, c)         \
    if (!(a) || !(b) || !(c)) \
        return SB200_E_POINTER;

int sb200_group_multiply_relinearize_host(sb200_group *g, size_t L, size_t batch, const uint64_t *a, const uint64_t *b, const sb200_group_key *key,
                                          uint64_t *out)
{
    NEED3(g, a, b)
    NEED3(key, out, g)
    if (key->key.size() != g->ctx.size())
        return SB200_E_INVALID_ARG;
    const size_t w = 2 * L * g->n;
    return fan_out(g, batch, [&](size_t i, size_t f, size_t c) {
        return sb200_multiply_relinearize_host(g->ctx[i], L, c, a + f * w, b + f * w, key->key[i], out + f * w);
    });
}

int sb200_group_relinearize_host(sb200_group *g, size_t L, size_t batch, const uint64_t *in3, const sb200_group_key *key, uint64_t *out2)
{
    NEED3(g, in3, key)
    NEED3(out2, g, g)
    if (key->key.size() != g->ctx.size())
        return SB200_E_INVALID_ARG;
    const size_t w = L * g->n;
    return fan_out(g, batch,
                   [&](size_t i, size_t f, size_t c) { return sb200_relinearize_host(g->ctx[i], L, c, in3 + f * 3 * w, key->key[i], out2 + f * 2 * w); });
}

int sb200_group_apply_galois_host(sb200_group *g, size_t L, size_t batch, const uint64_t *in2, uint32_t elt, const sb200_group_key *key, uint64_t *out2)
{
    NEED3(g, in2, key)
    NEED3(out2, g, g)
    if (key->key.size() != g->ctx.size())
        return SB200_E_INVALID_ARG;
    const size_t w = 2 * L * g->n;
    return fan_out(g, batch,
                   [&](size_t i, size_t f, size_t c) { return sb200_apply_galois_host(g->ctx[i], L, c, in2 + f * w, elt, key->key[i], out2 + f * w); });
}

int sb200_group_multiply_host(sb200_group *g, size_t L, size_t batch, const uint64_t *a, const uint64_t *b, uint64_t *out3)
{
    NEED3(g, a, b)
    NEED3(out3, g, g)
    const size_t w = L * g->n;
    return fan_out(g, batch,
                   [&](size_t i, size_t f, size_t c) { return sb200_multiply_host(g->ctx[i], L, c, a + f * 2 * w, b + f * 2 * w, out3 + f * 3 * w); });
}

int sb200_group_rescale_to_next_host(sb200_group *g, size_t L, size_t batch, const uint64_t *in2, uint64_t *out2)
{
    NEED3(g, in2, out2)
    if (L < 2)
        return SB200_E_INVALID_ARG;
    const size_t wi = 2 * L * g->n, wo = 2 * (L - 1) * g->n;
    return fan_out(g, batch, [&](size_t i, size_t f, size_t c) { return sb200_rescale_to_next_host(g->ctx[i], L, c, in2 + f * wi, out2 + f * wo); });
}

int sb200_group_mod_switch_to_next_host(sb200_group *g, size_t L, size_t batch, const uint64_t *in2, uint64_t *out2)
{
    NEED3(g, in2, out2)
    if (L < 2)
        return SB200_E_INVALID_ARG;
    const size_t wi = 2 * L * g->n, wo = 2 * (L - 1) * g->n;
    return fan_out(g, batch, [&](size_t i, size_t f, size_t c) { return sb200_mod_switch_to_next_host(g->ctx[i], L, c, in2 + f * wi, out2 + f * wo); });
}

int sb200_group_ntt_forward_host(sb200_group *g, size_t L, size_t size, size_t batch, uint64_t *h)
{
    NEED3(g, h, g)
    const size_t w = size * L * g->n;
    return fan_out(g, batch, [&](size_t i, size_t f, size_t c) { return sb200_ntt_forward_host(g->ctx[i], L, size, c, h + f * w); });
}

int sb200_group_ntt_inverse_host(sb200_group *g, size_t L, size_t size, size_t batch, uint64_t *h)
{
    NEED3(g, h, g)
    const size_t w = size * L * g->n;
    return fan_out(g, batch, [&](size_t i, size_t f, size_t c) { return sb200_ntt_inverse_host(g->ctx[i], L, size, c, h + f * w); });
}
} // extern "C"
