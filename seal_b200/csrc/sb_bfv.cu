// seal_b200/csrc/sb_bfv.cu -- BFV multiply (BEHZ): placeholder until the kernels land.
#include "sb_engine.cuh"
namespace sb
{
    struct BehzDev
    {
        sbh::BehzLevel host;
    };
    const sbh::BehzLevel &behz_host(Context &c, size_t L)
    {
        if (L < 1 || L > c.k)
            throw std::invalid_argument("no such level");
        auto it = c.behz.find(L);
        if (it == c.behz.end())
        {
            auto d = std::make_shared<BehzDev>();
            d->host = sbh::build_behz(c.n, c.q, L, c.t);
            it = c.behz.emplace(L, d).first;
        }
        return it->second->host;
    }
    void op_bfv_multiply(Context &, size_t, size_t, const u64 *, const u64 *, u64 *, cudaStream_t)
    {
        throw std::logic_error("BFV multiply: not implemented yet");
    }
} // namespace sb
