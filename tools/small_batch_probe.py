#!/usr/bin/env python
"""tools/small_batch_probe.py -- device-resident multiply+relinearize at small batches (the single-ciphertext calls of the C++ stand-in):
64-bit digit transforms vs the integer path, CKKS n = 32768, 16 primes.  Prints ms per call."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import seal_b200 as S  # noqa: E402

n, bits = 32768, [55] * 15 + [56]
mods = S.coeff_modulus_create(n, bits)
k, L = len(mods), len(mods) - 1
ctx = S.Context(S.CKKS, n, mods)
g = torch.Generator(device="cuda")
g.manual_seed(1)


def rand(prefix, rows):
    t = torch.empty((*prefix, rows, n), dtype=torch.int64, device="cuda")
    for i in range(rows):
        t[..., i, :] = torch.randint(0, mods[i], (*prefix, n), generator=g, dtype=torch.int64, device="cuda")
    return t


key = rand((L, 2), k)
rk = ctx.load_key(key.cpu().numpy().view("uint64"))
for B in (1, 2, 4, 8, 16, 64):
    a, b = rand((B, 2), L), rand((B, 2), L)
    out = torch.empty_like(a)
    for algo in (0, 2):
        ctx.set_limit(ctx.LIMIT_KS_ALGORITHM, algo)
        ctx.d_multiply_relinearize(a, b, rk, out, L, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.d_multiply_relinearize(a, b, rk, out, L, B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ctx.profile(True)
        ctx.d_multiply_relinearize(a, b, rk, out, L, B)
        torch.cuda.synchronize()
        prof = sorted(ctx.profile_read(), key=lambda r: -r[1])[:4]
        ctx.profile(False)
        print("B=%3d algo=%d  %.3f ms per call   top: %s" % (B, algo, ms, ", ".join("%s %.2f" % (p[0], p[1]) for p in prof)), flush=True)

# host-buffer entry points, one ciphertext per call (what seal_b200::Evaluator's single-ciphertext members use)
import time

import numpy as np

c3 = np.ascontiguousarray(rand((1, 3), L).cpu().numpy().view("uint64"))
for algo in (0, 2, 0, 2):
    ctx.set_limit(ctx.LIMIT_KS_ALGORITHM, algo)
    ctx.relinearize(c3, rk)
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.relinearize(c3, rk)
    print("host relinearize, 1 ciphertext, algo=%d: %.2f ms per call" % (algo, (time.perf_counter() - t0) * 100), flush=True)
