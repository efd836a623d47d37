// tests/cpp/host_probe.cpp -- exposes the product's host-side precomputation (seal_b200/csrc/sb_host.cpp, no CUDA) to
// the CPU test-suite so tables can be compared with the oracle / reference without a GPU.
#include "../../seal_b200/csrc/sb_host.hpp"
#include "../../seal_b200/csrc/sb_wire.hpp"
#include <cstring>
#include <stdexcept>

extern "C" {
int probe_tables(size_t n, unsigned long long q, unsigned long long *root, unsigned long long *rp, unsigned long long *rpq,
                 unsigned long long *irp, unsigned long long *inv_n, unsigned long long *fwd_w, unsigned long long *inv_w,
                 unsigned long long *ratio)
{
    try
    {
        sbh::PrimeTables t;
        t.build(n, q);
        *root = t.root;
        *inv_n = t.inv_n.w;
        ratio[0] = t.ratio_lo, ratio[1] = t.ratio_hi;
        for (size_t i = 0; i < n; i++)
        {
            rp[i] = t.root_powers[i].w, rpq[i] = t.root_powers[i].wq, irp[i] = t.inv_root_powers[i].w;
            fwd_w[i] = t.fwd[i].w, inv_w[i] = t.inv[i].w;
        }
        return 0;
    }
    catch (...)
    {
        return -1;
    }
}
size_t probe_bsk(size_t n, const unsigned long long *q, size_t L, unsigned long long t, unsigned long long *out)
{
    try
    {
        auto b = sbh::build_behz(n, std::vector<sbh::u64>(q, q + L), L, t);
        std::memcpy(out, b.Bsk.data(), b.nBsk * sizeof(unsigned long long));
        return b.nBsk;
    }
    catch (...)
    {
        return 0;
    }
}
int probe_galois_table(size_t n, unsigned elt, unsigned *out)
{
    auto t = sbh::galois_table_ntt(n, elt);
    std::memcpy(out, t.data(), n * sizeof(unsigned));
    return 0;
}
unsigned probe_elt_from_step(size_t n, int step)
{
    try
    {
        return sbh::galois_elt_from_step(n, step);
    }
    catch (...)
    {
        return 0;
    }
}
void probe_parms_id(int scheme, size_t n, const unsigned long long *q, size_t L, unsigned long long t, unsigned long long *out4)
{
    sbw::parms_id(scheme, n, q, L, t, out4);
}
void probe_blake2b_256(const void *in, size_t len, unsigned long long *out4) { sbw::blake2b_256(in, len, out4); }
// offsets of data()[index][j] inside a serialized KSwitchKeys object; returns the digit count or a negative status
long probe_kswitch_offsets(const unsigned char *stream, size_t len, size_t index, size_t *offsets, size_t capacity, size_t *L, size_t *n)
{
    try
    {
        sbw::KSwitchEntry e;
        sbw::inspect_kswitch(stream, len, index, e);
        if (e.offsets.size() > capacity)
            return -4;
        std::memcpy(offsets, e.offsets.data(), e.offsets.size() * sizeof(size_t));
        *L = e.L, *n = e.n;
        return static_cast<long>(e.offsets.size());
    }
    catch (const std::out_of_range &)
    {
        return -3;
    }
    catch (const std::invalid_argument &)
    {
        return -1;
    }
    catch (...)
    {
        return -2;
    }
}
int probe_batch_index_map(size_t n, unsigned *out)
{
    auto m = sbh::batch_index_map(n);
    std::memcpy(out, m.data(), n * sizeof(unsigned));
    return 0;
}
int probe_is_prime(unsigned long long v) { return sbh::is_prime(v) ? 1 : 0; }
}
