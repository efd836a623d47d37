#!/usr/bin/env python
"""tools/bench_ops.py -- device-resident throughput of every hot-path operation at the BASELINE.json config shapes
(uniform-random ciphertexts, synthetic keys), for profiles/r01_ops_table.md.  Not the contract benchmark (bench.py is)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import seal_b200 as S  # noqa: E402

CONFIGS = [
    ("cfg1 BFV n=4096 k=3 (BFVDefault)", S.BFV, 4096, None, 1024),
    ("cfg2 CKKS n=8192 k=4 (54-bit)", S.CKKS, 8192, [54] * 4, 1024),
    ("cfg3 CKKS n=32768 k=16 (55/56-bit)", S.CKKS, 32768, [55] * 15 + [56], 256),
    ("cfg4 BFV n=16384 k=8 (54-bit)", S.BFV, 16384, [54] * 8, 512),
    ("cfg5 CKKS n=65536 k=32 (55-bit)", S.CKKS, 65536, [55] * 32, 128),
]
BFV_DEFAULT_4096 = [0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001]


def rand_rows(mods, n, prefix, nprimes, g):
    t = torch.empty((*prefix, nprimes, n), dtype=torch.int64, device="cuda")
    for i in range(nprimes):
        t[..., i, :] = torch.randint(0, mods[i], (*prefix, n), generator=g, dtype=torch.int64, device="cuda")
    return t


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    rows = []
    for name, scheme, n, bits, batch in CONFIGS:
        mods = BFV_DEFAULT_4096 if bits is None else S.coeff_modulus_create(n, bits)
        k, L = len(mods), len(mods) - 1
        t = 1032193 if scheme == S.BFV and n == 4096 else (786433 if scheme == S.BFV else 0)
        ctx = S.Context(scheme, n, mods, t)
        key = ctx.load_key(rand_rows(mods, n, (L, 2), k, g).cpu().numpy().view(np.uint64))
        a, b = rand_rows(mods, n, (batch, 2), L, g), rand_rows(mods, n, (batch, 2), L, g)
        o3 = torch.empty((batch, 3, L, n), dtype=torch.int64, device="cuda")
        o2 = torch.empty((batch, 2, L, n), dtype=torch.int64, device="cuda")
        om = torch.empty((batch, 2, max(L - 1, 1), n), dtype=torch.int64, device="cuda")
        res = {"config": name, "batch": batch}
        res["multiply"] = batch / timeit(lambda: ctx.d_multiply(a, b, o3, L, batch)) * 1e3
        res["relinearize"] = batch / timeit(lambda: ctx.d_relinearize(o3, key, o2, L, batch)) * 1e3
        res["multiply+relinearize"] = batch / timeit(lambda: ctx.d_multiply_relinearize(a, b, key, o2, L, batch)) * 1e3
        ms = timeit(lambda: ctx.d_multiply_plain(a, b, o2, L, 2, batch))  # b's first [batch][L][n] words serve as the plaintexts
        res["multiply_plain"] = batch / ms * 1e3
        res["multiply_plain_GBps"] = batch * 5 * L * n * 8 / (ms * 1e-3) / 1e9
        elt = ctx.galois_elt_from_step(1)
        res["rotate (1 step)"] = batch / timeit(lambda: ctx.d_apply_galois(a, elt, key, o2, L, batch)) * 1e3
        if L > 1:
            f = ctx.d_rescale_to_next if scheme == S.CKKS else ctx.d_mod_switch_to_next
            res["rescale/mod_switch"] = batch / timeit(lambda: f(a, om, L, batch)) * 1e3
        ms = timeit(lambda: (ctx.d_ntt_inverse(a, L, 2, batch), ctx.d_ntt_forward(a, L, 2, batch)))
        res["ct NTT+INTT"] = batch / ms * 1e3
        res["ntt_GBps"] = 2 * batch * 2 * L * 2 * n * 8 / (ms * 1e-3) / 1e9
        rows.append(res)
        print(json.dumps(res), flush=True)
        del ctx, key, a, b, o3, o2, om
        torch.cuda.empty_cache()
    return rows


# the reference on ONE core (survey container), ops/s, for scale only
CPU_1CORE = {
    "cfg1": {"multiply": 277.8, "relinearize": 1041.7, "multiply+relinearize": 219.3, "rotate (1 step)": 1408.5, "ct NTT+INTT": 3225.8},
    "cfg2": {"multiply": 2083.3, "relinearize": 416.7, "multiply+relinearize": 347.2, "rotate (1 step)": 434.8,
             "rescale/mod_switch": 1041.7, "ct NTT+INTT": 833.3},
    "cfg3": {"multiply": 82.6, "relinearize": 5.8, "multiply+relinearize": 5.5, "rotate (1 step)": 4.4, "rescale/mod_switch": 37.7,
             "ct NTT+INTT": 21.8},
    "cfg4": {"multiply": 19.3, "relinearize": 46.9, "multiply+relinearize": 13.7, "rotate (1 step)": 48.3, "ct NTT+INTT": 163.9},
    "cfg5": {"multiply": 21.0, "relinearize": 0.7, "multiply+relinearize": 0.7, "rotate (1 step)": 0.8, "rescale/mod_switch": 8.1,
             "ct NTT+INTT": 7.4},
}
HBM_PEAK_GBPS = 6571.6  # MEASURED_PEAKS.json


def table(rows):
    """markdown for profiles/r01_ops_table.md from the JSON lines main() prints"""
    out = ["# Round 1: device-resident throughput of every hot-path op at the BASELINE.json config shapes", "",
           "`python tools/bench_ops.py` on one B200 (CUDA events, 3 repetitions after a warm-up, uniform-random ciphertexts and "
           "synthetic keys resident in HBM); table written by `python tools/bench_ops.py --table <json lines>`.",
           "`cpu x1` = the reference on ONE core as measured in BASELINE.md section 2 (survey container), for scale only; "
           "the contract benchmark is bench.py.", "",
           "| config | batch | op | B200 ops/s | reference 1-core ops/s | ratio |", "|---|---|---|---|---|---|"]
    for r in rows:
        cpu = CPU_1CORE.get(r["config"][:4], {})
        for op, v in r.items():
            if op in ("config", "batch"):
                continue
            if op.endswith("_GBps"):
                what = "NTT algorithmic GB/s (2*n*8 B per row per transform)" if op == "ntt_GBps" else \
                    "multiply_plain algorithmic GB/s (5*L*n*8 B per size-2 ciphertext)"
                out.append(f"| {r['config']} | {r['batch']} | {what} | {v:.0f} | | {v / HBM_PEAK_GBPS:.2f} of HBM peak |")
            else:
                c = cpu.get(op)
                out.append(f"| {r['config']} | {r['batch']} | {op} | {v:.0f} | {c if c else '-'} | {f'{v / c:.0f}x' if c else '-'} |")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--table":
        sys.stdout.write(table([json.loads(ln) for ln in open(sys.argv[2]) if ln.startswith("{")]))
    else:
        main()
