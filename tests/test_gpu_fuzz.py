"""Randomised parity sweep (-m gpu): random ring sizes, prime counts / bit sizes (crossing the 57-bit guard-free boundary
and the 4x digit-reduction boundary), levels and batch sizes, all three schemes, against the plain-C oracle."""
import numpy as np
import pytest

import oracle as O
from common import rand_ct

pytestmark = pytest.mark.gpu


def sb():
    import seal_b200

    return seal_b200


def _cases():
    rng = np.random.default_rng(0xB200)
    out = []
    for i in range(28):
        logn = int(rng.choice([4, 6, 8, 10, 11, 12, 12, 13, 13, 14]))
        k = int(rng.integers(2, 6))
        lo = max(logn + 4, 20)
        bits = [int(rng.integers(lo, 61)) for _ in range(k)]
        if i % 4 == 0:
            bits = [int(rng.choice([56, 57, 58]))] * k  # around the guard-free threshold
        if i % 7 == 0:
            bits[0], bits[-1] = 60, lo  # large spread: digit re-reduction must stay on
        scheme = "bfv" if i % 3 == 0 else "ckks"
        out.append((logn, tuple(bits), scheme, int(rng.integers(1, 5)), i))
    for i in range(28, 38):  # BGV; odd seeds use a plain modulus above the small primes
        logn = int(rng.choice([4, 8, 11, 12, 13, 14]))
        k = int(rng.integers(2, 6))
        lo = max(logn + 4, 20)
        bits = [int(rng.integers(lo, 61)) for _ in range(k)]
        if i % 5 == 0:
            bits[0], bits[-1] = 60, lo
        out.append((logn, tuple(bits), "bgv", int(rng.integers(1, 5)), i))
    return out


@pytest.mark.parametrize("logn,bits,scheme,batch,seed", _cases())
def test_random_config_vs_oracle(logn, bits, scheme, batch, seed):
    n = 1 << logn
    try:
        mods = O.coeff_modulus_create(n, list(bits))
    except RuntimeError:
        pytest.skip("not enough primes of that size")
    if len(set(mods)) != len(mods):
        pytest.skip("duplicate primes")
    k = len(mods)
    t = 65537 if scheme == "bfv" else 0
    sid = sb().BFV if scheme == "bfv" else sb().CKKS
    if scheme == "bgv":
        t, sid = (2147483647 if seed % 2 else 65537), sb().BGV
    ctx = sb().Context(sid, n, mods, t)
    oc = O.Oracle(sid, n, mods, t)
    rng = np.random.default_rng(seed)
    L = int(rng.integers(1, k))
    key = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(k)]) for _ in range(2)])
                    for _ in range(k - 1)])
    rk = ctx.load_key(key)
    if logn >= 12:
        # these shapes have fewer than 6 digits (automatic mode = the 64-bit digit transforms): every other case runs key switching
        # through the exact integer convolution on the auxiliary primes instead (seal_b200/csrc/sb_ksint.cu) -- mixed prime sizes,
        # levels below the key level, Galois views, all three schemes
        ctx.set_limit(sb().Context.LIMIT_KS_ALGORITHM, 2 if seed % 2 else 0)
    a, b = rand_ct(rng, mods, n, 2, L, batch), rand_ct(rng, mods, n, 2, L, batch)
    i = batch - 1
    got = ctx.multiply_relinearize(a, b, rk)
    assert (got[i] == oc.multiply_relin(L, a[i], b[i], key)).all()
    assert (ctx.transform_from_ntt(ctx.transform_to_ntt(a)) == a).all()
    if n >= 4:
        step = int(rng.integers(1, max(2, n // 2)))
        e = O.galois_elt_from_step(n, step if rng.integers(0, 2) else -step)
        assert (ctx.apply_galois(a, e, rk)[i] == oc.apply_galois(L, a[i], e, key)).all()
    if L > 1:
        if scheme == "ckks":
            assert (ctx.rescale_to_next(a)[i] == oc.rescale(L, a[i])).all()
        elif scheme == "bgv":
            assert (ctx.mod_switch_to_next(a)[i] == oc.bgv_mod_switch(L, a[i])).all()
        else:
            assert (ctx.mod_switch_to_next(a)[i] == oc.bfv_mod_switch(L, a[i])).all()
