"""The integer key-switching path (-m gpu; seal_b200/csrc/sb_ksint.cu): its 32-bit transforms modulo the auxiliary primes are
checked by what a negacyclic NTT must satisfy (round trip, convolution theorem against a direct O(n^2) product), and the whole
path is compared with the 64-bit digit-transform path of the same library and with the oracle / the reference on the same
inputs (switch_key_inplace, evaluator.cpp:2561-2867) -- both must give the reference's words."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def sb():
    import seal_b200

    return seal_b200


def negacyclic(a, b, p):
    """direct product in Z_p[x]/(x^n + 1); a, b < p < 2^29"""
    n = len(a)
    a, b = a.astype(np.uint64), b.astype(np.uint64)
    acc = np.zeros(n, dtype=np.uint64)
    for i in range(n):
        # x^i * b: coefficients wrap with a sign
        rot = np.concatenate([(p - b[n - i:]) % p, b[:n - i]]) if i else b
        acc = (acc + a[i] * rot) % p
    return acc


@pytest.mark.parametrize("n", [4096, 8192, 65536])
def test_small_transforms_roundtrip_and_convolution(n):
    S = sb()
    mods = O.coeff_modulus_create(n, [50, 50, 51])
    ctx = S.Context(S.CKKS, n, mods)
    primes = ctx.ksint_primes()
    assert len(primes) >= 4 and all(p < 2 ** 29 and p % (2 * n) == 1 for p in primes)
    rng = np.random.default_rng(n)
    rows = 11  # not a multiple of the rows a CTA loops over
    x = rng.integers(0, 2 ** 61, (rows, n), dtype=np.uint64)
    x[0, :] = 0
    x[0, 1] = 1  # the monomial x: its transform is the table of odd psi powers, all nonzero
    f = ctx.ksint_forward(x)  # [S][rows][n]
    assert f.shape == (len(primes), rows, n)
    for t, p in enumerate(primes):
        assert (f[t] < p).all(), "forward outputs must be canonical"
    back = ctx.ksint_inverse(np.ascontiguousarray(f.transpose(1, 0, 2)))  # [rows][S][n]
    for t, p in enumerate(primes):
        assert (back[:, t, :] < 2 * p).all()
        assert ((back[:, t, :].astype(np.uint64) % p) == ((x % p) * n) % p).all(), f"round trip, prime {p}"
    if n <= 8192:
        a, b = x[1], x[2]
        fa, fb = f[:, 1, :].astype(np.uint64), f[:, 2, :].astype(np.uint64)
        prod = np.stack([(fa[t] * fb[t]) % p for t, p in enumerate(primes)])[None].astype(np.uint32)  # [1][S][n]
        conv = ctx.ksint_inverse(prod)[0]
        for t, p in enumerate(primes):
            want = negacyclic(a % p, b % p, p)
            assert ((conv[t].astype(np.uint64) % p) == (want * n) % p).all(), f"convolution theorem, prime {p}"


def _keyswitch_case(scheme_name, n, bits, batch, t=0):
    S = sb()
    scheme = getattr(S, scheme_name)
    mods = O.coeff_modulus_create(n, bits)
    k, L = len(mods), len(mods) - 1
    ctx = S.Context(scheme, n, mods, t) if t else S.Context(scheme, n, mods)
    oc = O.Oracle(getattr(O, scheme_name), n, mods, t) if t else O.Oracle(getattr(O, scheme_name), n, mods)
    rng = np.random.default_rng(7 * n + k)
    key = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(k)]) for _ in range(2)])
                    for _ in range(L)])
    c3 = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(3)])
                   for _ in range(batch)])
    rk = ctx.load_key(key)
    assert ctx.ksint_primes(), "the integer path must be available at n >= 4096"
    ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 2)
    got_int = ctx.relinearize(c3, rk)
    ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 0)
    got_64 = ctx.relinearize(c3, rk)
    for b in range(batch):
        want = oc.relinearize(L, c3[b], key)
        assert (got_64[b] == want).all(), "64-bit digit-transform path vs oracle"
        assert (got_int[b] == want).all(), "integer path vs oracle"
    # a lower level uses a subset of the digits and of the output primes (evaluator.cpp:2617-2640)
    if L >= 2:
        ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 2)
        low = np.ascontiguousarray(c3[:, :, :L - 1, :])
        got = ctx.relinearize(low, rk)
        for b in range(batch):
            assert (got[b] == oc.relinearize(L - 1, low[b], key)).all(), "integer path, lower level"


@pytest.mark.parametrize("n,bits,batch", [(4096, [36, 36, 37], 3), (8192, [50, 50, 50, 51], 5), (16384, [60, 60, 60], 2),
                                          (32768, [55] * 5, 9), (131072, [58, 58, 59], 1)])
def test_relinearize_integer_path_vs_oracle_ckks(n, bits, batch):
    _keyswitch_case("CKKS", n, bits, batch)


def test_relinearize_integer_path_vs_oracle_bfv():
    _keyswitch_case("BFV", 4096, [36, 36, 37], 3, t=65537)


def test_relinearize_integer_path_vs_oracle_bgv():
    _keyswitch_case("BGV", 8192, [50, 50, 50, 51], 3, t=65537)


@pytest.mark.parametrize("scheme_name,n,bits,t", [("CKKS", 8192, [50, 50, 50, 51], 0), ("BFV", 4096, [36, 36, 37], 65537),
                                                 ("BGV", 4096, [40, 40, 40], 65537), ("CKKS", 4096, [30] * 9, 0)])
def test_rotation_and_fused_multiply_integer_path_vs_oracle(scheme_name, n, bits, t):
    """apply_galois (Galois views of the target: NTT-form permutation for CKKS / BGV, coefficient-form automorphism with sign flips
    for BFV, evaluator.cpp:2384-2502) and multiply + relinearize through the integer path, every word against the oracle; the last
    case has 8 digits (the automatic mode's territory) of 30-bit primes (4 auxiliary primes)"""
    S = sb()
    scheme = getattr(S, scheme_name)
    mods = O.coeff_modulus_create(n, bits)
    k, L = len(mods), len(mods) - 1
    ctx = S.Context(scheme, n, mods, t) if t else S.Context(scheme, n, mods)
    oc = O.Oracle(getattr(O, scheme_name), n, mods, t) if t else O.Oracle(getattr(O, scheme_name), n, mods)
    rng = np.random.default_rng(3 * n + k)
    key = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(k)]) for _ in range(2)])
                    for _ in range(L)])
    rk = ctx.load_key(key)
    batch = 3
    c2 = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(2)])
                   for _ in range(batch)])
    ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 2)
    for elt in (3, 2 * n - 1, O.galois_elt_from_step(n, 1)):
        got = ctx.apply_galois(c2, elt, rk)
        for b in range(batch):
            assert (got[b] == oc.apply_galois(L, c2[b], elt, key)).all(), ("apply_galois", elt, b)
    if scheme_name == "CKKS":
        d2 = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(2)])
                       for _ in range(batch)])
        got = ctx.multiply_relinearize(c2, d2, rk)
        for b in range(batch):
            assert (got[b] == oc.multiply_relin(L, c2[b], d2[b], key)).all(), ("multiply_relinearize", b)
        # automatic mode picks the same words whatever path a level runs
        ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 1)
        assert (ctx.multiply_relinearize(c2, d2, rk) == got).all()


@pytest.mark.parametrize("n,bits,batch", [(4096, [30] * 9, 37), (4096, [30] * 42, 18)])
def test_product_kernel_shapes_vs_oracle(n, bits, batch):
    """the two product kernels of the integer path: the key-tile kernel walks over 16 ciphertexts per iteration (batch 37: two full
    iterations and a ragged one), and falls back to the register-tile kernel when the key tile of a level does not fit shared memory
    (41 digits); sampled ciphertexts against the oracle, all of them against the 64-bit path"""
    S = sb()
    mods = O.coeff_modulus_create(n, bits)
    k, L = len(mods), len(mods) - 1
    ctx = S.Context(S.CKKS, n, mods)
    oc = O.Oracle(O.CKKS, n, mods)
    rng = np.random.default_rng(n + k + batch)
    key = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(k)]) for _ in range(2)])
                    for _ in range(L)])
    c3 = np.stack([np.stack([np.stack([rng.integers(0, mods[i], n, dtype=np.uint64) for i in range(L)]) for _ in range(3)])
                   for _ in range(batch)])
    rk = ctx.load_key(key)
    ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 2)
    got = ctx.relinearize(c3, rk)
    ctx.set_limit(S.Context.LIMIT_KS_ALGORITHM, 0)
    assert (ctx.relinearize(c3, rk) == got).all()
    for b in (0, 15, 16, batch - 1):
        assert (got[b] == oc.relinearize(L, c3[b], key)).all(), b
