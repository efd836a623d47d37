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


if __name__ == "__main__":
    main()
