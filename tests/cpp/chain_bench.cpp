// tests/cpp/chain_bench.cpp -- end-to-end figures measured THROUGH THE C++ DROP-IN CLASS (seal_b200::Evaluator on seal::Ciphertext /
// seal_b200::CiphertextBatch), for bench.py's "configs.cpp" entry.  BASELINE.json configs[2] shape: CKKS n=32768, 16 primes
// (CoeffModulus::BFVDefault(32768)), batch 256, chain  a <- rescale(relin(a*b)); b <- mod_switch_to_next(b)  of depth 8 -- the usage
// pattern of native/tests/seal/evaluator.cpp:3513-3780.
//   chain_e2e      upload(std::vector<Ciphertext>) -> 8 levels on the device -> download: wall clock, host objects in and out
//   chain_e2e_pipelined  the same work cut into 4 slices: one host thread uploads slice k+1 and another downloads slice k-1 while the
//                  device runs the chain of slice k (upload / download run on their own streams inside the library)
//   chain_device   the same 8 levels with the batch already resident (upload / download outside the timed region)
//   single_calls   the reference-signature members on one seal::Ciphertext at a time (every call moves its operands both ways)
// One ciphertext of the batch is checked against the reference's own seal::Evaluator (linked from oracle/_ref/libseal.so).
// TEST / BENCH INFRASTRUCTURE: links the reference; built only where /root/reference exists; the binary travels to the GPU box.
#include "seal_b200/evaluator.hpp"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <future>
#include <random>
#include <thread>

using namespace seal;
using clk = std::chrono::steady_clock;

static double secs(clk::time_point a, clk::time_point b)
{
    return std::chrono::duration<double>(b - a).count();
}

static bool same_ct(const Ciphertext &a, const Ciphertext &b)
{
    if (a.parms_id() != b.parms_id() || a.size() != b.size() || a.is_ntt_form() != b.is_ntt_form())
        return false;
    double sa = a.scale(), sb = b.scale();
    if (std::memcmp(&sa, &sb, sizeof(double)) != 0)
        return false;
    return std::memcmp(a.data(), b.data(), a.size() * a.coeff_modulus_size() * a.poly_modulus_degree() * 8) == 0;
}

int main(int argc, char **argv)
{
    try
    {
        const size_t n = 32768, batch = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 256, depth = 8;
        EncryptionParameters parms(scheme_type::ckks);
        parms.set_poly_modulus_degree(n);
        parms.set_coeff_modulus(CoeffModulus::BFVDefault(n)); // 15 x 55-bit + 56-bit (util/globals.cpp:66-72)
        SEALContext context(parms, true, sec_level_type::none);
        KeyGenerator keygen(context);
        RelinKeys rlk;
        keygen.create_relin_keys(rlk);
        seal::Evaluator ref(context);
        seal_b200::Evaluator gpu(context);

        // synthetic ciphertexts as native/bench makes them (bench.h:195-270): uniform residues, size 2, NTT form
        auto cd = context.first_context_data();
        const auto &q = cd->parms().coeff_modulus();
        const size_t L = q.size();
        const double scale = std::pow(2.0, 54);
        std::mt19937_64 rng(0x5EA1);
        auto make = [&](Ciphertext &c) {
            c.resize(context, cd->parms_id(), 2);
            c.is_ntt_form() = true;
            c.scale() = scale;
            for (size_t p = 0; p < 2; p++)
                for (size_t i = 0; i < L; i++)
                    for (size_t j = 0; j < n; j++)
                        c.data(p)[i * n + j] = rng() % q[i].value();
        };
        std::vector<Ciphertext> a(batch), b(batch);
        for (size_t i = 0; i < batch; i++)
            make(a[i]), make(b[i]);

        auto run_chain = [&](seal_b200::CiphertextBatch &da, seal_b200::CiphertextBatch &db) {
            for (size_t d = 0; d < depth; d++)
            {
                gpu.multiply_relinearize_inplace(da, db, rlk);
                gpu.rescale_to_next_inplace(da);
                gpu.mod_switch_to_next_inplace(db);
                da.scale() = scale; // synthetic data: keep the scale where it started so that all 8 levels stay in bounds
                db.scale() = scale;
            }
        };
        seal_b200::CiphertextBatch da, db;
        std::vector<Ciphertext> out;
        // warm-up: key upload, slab growth, table use
        gpu.upload(a, da), gpu.upload(b, db);
        run_chain(da, db);
        gpu.download(da, out);

        // device-resident chain
        gpu.upload(a, da), gpu.upload(b, db);
        gpu.synchronize();
        auto t0 = clk::now();
        run_chain(da, db);
        gpu.synchronize();
        const double dev_s = secs(t0, clk::now());

        // end to end: host objects in, host objects out
        t0 = clk::now();
        gpu.upload(a, da), gpu.upload(b, db);
        run_chain(da, db);
        gpu.download(da, out);
        const double e2e_s = secs(t0, clk::now());

        // pipelined end to end: slices of the batch flow upload -> chain -> download on three host threads
        const size_t slices = std::min<size_t>(4, batch), per = (batch + slices - 1) / slices;
        std::vector<seal_b200::CiphertextBatch> sa(slices), sb(slices);
        std::vector<Ciphertext> pout(batch);
        double pipe_s = 0;
        for (int rep = 0; rep < 2; rep++) // the first repetition sizes the slabs of every slice
        {
            std::vector<std::promise<void>> up(slices), enq(slices);
            t0 = clk::now();
            std::thread uploader([&] {
                for (size_t k = 0; k < slices; k++)
                {
                    const size_t f = k * per, c = std::min(per, batch - f);
                    gpu.upload(a.data() + f, c, sa[k]);
                    gpu.upload(b.data() + f, c, sb[k]);
                    up[k].set_value();
                }
            });
            std::thread downloader([&] {
                for (size_t k = 0; k < slices; k++)
                {
                    enq[k].get_future().wait();
                    gpu.download(sa[k], pout.data() + k * per);
                }
            });
            for (size_t k = 0; k < slices; k++)
            {
                up[k].get_future().wait();
                run_chain(sa[k], sb[k]);
                enq[k].set_value();
            }
            uploader.join();
            downloader.join();
            pipe_s = secs(t0, clk::now());
        }
        bool pipe_ok = true;
        for (size_t i = 0; i < batch; i += std::max<size_t>(1, batch / 8))
            pipe_ok = pipe_ok && same_ct(out[i], pout[i]);

        // check one ciphertext of the batch against the reference evaluator (same chain)
        const size_t pick = batch - 1;
        Ciphertext ra = a[pick], rb = b[pick];
        t0 = clk::now();
        for (size_t d = 0; d < depth; d++)
        {
            ref.multiply_inplace(ra, rb);
            ref.relinearize_inplace(ra, rlk);
            ref.rescale_to_next_inplace(ra);
            ref.mod_switch_to_next_inplace(rb);
            ra.scale() = scale, rb.scale() = scale;
        }
        const double ref_chain_s = secs(t0, clk::now());
        const bool ok = same_ct(ra, out[pick]) && same_ct(ra, pout[pick]) && pipe_ok;

        // the reference-signature members, one seal::Ciphertext per call (first level only): what a user gets without batches
        const size_t singles = std::min<size_t>(batch, 16);
        t0 = clk::now();
        for (size_t i = 0; i < singles; i++)
        {
            Ciphertext x = a[i];
            gpu.multiply_inplace(x, b[i]);
            gpu.relinearize_inplace(x, rlk);
            gpu.rescale_to_next_inplace(x);
        }
        const double single_s = secs(t0, clk::now()) / singles;

        const double h2d = 2.0 * batch * 2 * L * n * 8, d2h = 1.0 * batch * 2 * (L - depth) * n * 8;
        std::printf("{\"harness\": \"tests/cpp/chain_bench.cpp (seal_b200::Evaluator + CiphertextBatch, C++)\", "
                    "\"config\": \"CKKS n=32768, 16 primes, batch %zu, depth-%zu chain\", "
                    "\"chain_e2e\": {\"value\": %.1f, \"unit\": \"chain steps/s\", \"seconds\": %.4f, \"h2d_bytes\": %.0f, \"d2h_bytes\": %.0f}, "
                    "\"chain_e2e_pipelined\": {\"value\": %.1f, \"unit\": \"chain steps/s\", \"seconds\": %.4f, \"slices\": %zu}, "
                    "\"chain_device\": {\"value\": %.1f, \"unit\": \"chain steps/s\", \"seconds\": %.4f}, "
                    "\"e2e_over_device\": %.3f, \"pipelined_e2e_over_device\": %.3f, "
                    "\"single_ciphertext_calls\": {\"value\": %.1f, \"unit\": \"multiply+relinearize+rescale/s, one seal::Ciphertext per call\"}, "
                    "\"reference_cpu_one_thread\": {\"value\": %.2f, \"unit\": \"chain steps/s\"}, "
                    "\"verified\": {\"index\": %zu, \"ok\": %s, \"against\": \"seal::Evaluator of the reference, same chain\"}}\n",
                    batch, depth, batch * depth / e2e_s, e2e_s, h2d, d2h, batch * depth / pipe_s, pipe_s, slices, batch * depth / dev_s, dev_s,
                    dev_s / e2e_s, dev_s / pipe_s, 1.0 / single_s,
                    depth / ref_chain_s, pick, ok ? "true" : "false");
        return ok ? 0 : 1;
    }
    catch (const std::exception &e)
    {
        std::printf("{\"unavailable\": \"%s\"}\n", e.what());
        return 2;
    }
}
