"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libsealref.so, i.e. microsoft/SEAL 4.4.3 compiled
from /root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin both the oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py) on boxes where the
reference library is not available.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import refseal as R  # noqa: E402


def rnd(rng, mods, n, size, L):
    return np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(size)])


def make(name, scheme, n, bits, t_bits, seed):
    rng = np.random.default_rng(seed)
    mods = R.coeff_modulus_create(n, bits)
    t = R.plain_modulus_batching(n, t_bits) if scheme != R.CKKS else 0
    rc = R.RefContext(scheme, n, mods, t, seed=0x5EA1)
    k = len(mods)
    out = dict(scheme=scheme, n=n, moduli=np.array(mods, dtype=np.uint64), t=np.uint64(t))
    out["roots"] = np.array([rc.ntt_root(i) for i in range(k)], dtype=np.uint64)
    out["relin_key"] = rc.relin_key()
    elts = [3, 2 * n - 1, rc.galois_elt_from_step(-2)]
    out["galois_elts"] = np.array(elts, dtype=np.uint32)
    for e in elts:
        out[f"galois_key_{e}"] = rc.galois_key(e)
    for L in range(k - 1, 0, -1):
        a, b = rnd(rng, mods, n, 2, L), rnd(rng, mods, n, 2, L)
        out[f"L{L}_a"], out[f"L{L}_b"] = a, b
        out[f"L{L}_ntt_fwd_a"] = rc.ntt_forward(L, a)
        out[f"L{L}_ntt_inv_a"] = rc.ntt_inverse(L, a)
        m = rc.multiply(L, a, b)
        out[f"L{L}_mul"] = m
        out[f"L{L}_relin"] = rc.relinearize(L, m)
        if L > 1:
            out[f"L{L}_modswitch_a"] = rc.rescale(L, a) if scheme == R.CKKS else rc.mod_switch(L, a)
        for e in elts:
            out[f"L{L}_galois_{e}"] = rc.apply_galois(L, a, e)
        if scheme == R.BFV:
            out[f"L{L}_bsk"] = np.array(rc.base_bsk(L), dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k_: getattr(v, "shape", None) for k_, v in list(out.items())[:6]})


if __name__ == "__main__":
    make("ckks_n128", R.CKKS, 128, [40, 30, 35, 41], 0, 11)
    make("bfv_n128", R.BFV, 128, [36, 36, 37], 17, 12)
    make("ckks_n1024", R.CKKS, 1024, [50, 50, 50], 0, 13)
    make("bgv_n128", R.BGV, 128, [36, 36, 37, 38], 17, 14)
