"""Shared helpers for the test-suite: golden fixtures, deterministic synthetic ciphertexts."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rand_ct(rng, moduli, n, size, L, batch=None):
    """uniform residues in [0, q_i), laid out [size][L][n] (or [batch][size][L][n]) like seal::Ciphertext::data()"""
    def one():
        return np.stack(
            [np.stack([rng.integers(0, moduli[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(size)])
    if batch is None:
        return one()
    return np.stack([one() for _ in range(batch)])
