"""Replays the reference's in-source known-answer tests for the RNS base-conversion steps against the oracle:
native/tests/seal/util/rns.cpp -- BaseConverterTest.Convert/ConvertArray (:347-438), RNSToolTest.FastBConvMTilde
(:460-537), MontgomeryReduction (:539-672), FastFloor (:674-787), FastBConvSK (:789-853),
DivideAndRoundQLastInplace (:904-1011).  Arrays are [rns component][coefficient] like the reference's RNSIter."""
import numpy as np

import oracle as O

MT = 1 << 32  # m_tilde, rns.cpp:635


def u(rows):
    return np.array(rows, dtype=np.uint64)


def test_base_converter_convert():
    # rns.cpp:347-400 (one value per base element = arrays of length 1)
    cases = [([2], [2], [[0], [0]]), ([2], [2], [[1], [1]]), ([2], [3], [[1], [1]]), ([3], [2], [[2], [0]]),
             ([2, 3], [2], [[1, 1], [1]]), ([2, 3], [2], [[0, 2], [0]]), ([2, 3], [2], [[1, 0], [1]]),
             ([2, 3], [2, 3], [[1, 2], [1, 2]]), ([2, 3], [2, 3], [[0, 2], [0, 2]]),
             ([2, 3], [3, 4, 5], [[0, 0], [0, 0, 0]]), ([2, 3], [3, 4, 5], [[1, 1], [1, 3, 2]]),
             ([2, 3], [3, 4, 5], [[1, 2], [2, 1, 0]]), ([3, 4, 5], [2, 3], [[1, 1, 1], [1, 1]])]
    for ib, ob, (x, want) in cases:
        got = O.fastbconv_array(ib, ob, u([[v] for v in x]))
        assert [int(r[0]) for r in got] == want, (ib, ob, x)


def test_base_converter_convert_array():
    # rns.cpp:402-438: in/out stored modulus-major ("array-major" in the test's words): [component][3 coefficients]
    got = O.fastbconv_array([3], [2], u([[0, 1, 2]]))
    assert got.tolist() == [[0, 1, 0]]
    got = O.fastbconv_array([2, 3], [2], u([[0, 1, 0], [0, 1, 2]]))
    assert got.tolist() == [[0, 1, 0]]
    got = O.fastbconv_array([2, 3], [2, 3], u([[1, 1, 0], [1, 2, 2]]))
    assert got.tolist() == [[1, 1, 0], [1, 2, 2]]
    got = O.fastbconv_array([2, 3], [3, 4, 5], u([[0, 1, 1], [0, 1, 2]]))
    assert got.tolist() == [[0, 1, 2], [0, 3, 1], [0, 2, 0]]


def test_fastbconv_m_tilde():
    # rns.cpp:460-537
    n = 2
    bsk = O.behz_base(n, [3])
    base = bsk + [MT]
    assert (O.behz_fastbconv_m_tilde(n, [3], u([[0, 0]])) == 0).all()
    out = O.behz_fastbconv_m_tilde(n, [3], u([[1, 2]]))
    t1, t2 = MT % 3, (2 * MT) % 3
    for s, p in enumerate(base):
        assert int(out[s, 0]) == t1 % p and int(out[s, 1]) == t2 % p
    bsk = O.behz_base(n, [3, 5])
    base = bsk + [MT]
    assert len(base) == 4
    out = O.behz_fastbconv_m_tilde(n, [3, 5], u([[1, 1], [2, 2]]))
    temp = ((2 * MT) % 3) * 5 + ((4 * MT) % 5) * 3
    for s, p in enumerate(base):
        assert int(out[s, 0]) == temp % p and int(out[s, 1]) == temp % p


def test_montgomery_reduction_sm_mrq():
    # rns.cpp:539-672
    n = 2
    assert (O.behz_sm_mrq(n, [3], u([[0, 0]] * 3)) == 0).all()
    out = O.behz_sm_mrq(n, [3], u([[MT, 2 * MT], [MT, 2 * MT], [0, 0]]))
    assert out.tolist() == [[1, 2], [1, 2]]
    assert (O.behz_sm_mrq(n, [3], u([[3, 3]] * 3)) == 0).all()
    out = O.behz_sm_mrq(n, [3, 5], u([[MT, 2 * MT]] * 3 + [[0, 0]]))
    assert out.tolist() == [[1, 2]] * 3
    assert (O.behz_sm_mrq(n, [3, 5], u([[15, 30]] * 4)) == 0).all()
    out = O.behz_sm_mrq(n, [3, 5], u([[2 * MT + 15, 2 * MT + 30]] * 4))
    assert (out == 2).all()


def test_fast_floor():
    # rns.cpp:674-787
    n = 2
    assert (O.behz_fast_floor(n, [3], u([[0, 0]] * 3)) == 0).all()
    assert O.behz_fast_floor(n, [3], u([[15, 3]] * 3)).tolist() == [[5, 1], [5, 1]]
    assert O.behz_fast_floor(n, [3], u([[17, 4]] * 3)).tolist() == [[5, 1], [5, 1]]
    assert O.behz_fast_floor(n, [3, 5], u([[15, 30]] * 5)).tolist() == [[1, 2]] * 3
    out = O.behz_fast_floor(n, [3, 5], u([[21, 32]] * 5))  # the reference asserts |result - floor| <= 1 here
    assert all(abs(int(r[0]) - 1) <= 1 and abs(int(r[1]) - 2) <= 1 for r in out)


def test_fastbconv_sk():
    # rns.cpp:789-853
    n = 2
    assert (O.behz_fastbconv_sk(n, [3], u([[0, 0]] * 2)) == 0).all()
    assert O.behz_fastbconv_sk(n, [3], u([[1, 2]] * 2)).tolist() == [[1, 2]]
    assert O.behz_fastbconv_sk(n, [3, 5], u([[1, 2]] * 3)).tolist() == [[1, 2], [1, 2]]


def test_divide_and_round_q_last():
    # rns.cpp:904-1011
    assert O.divide_and_round_q_last([13, 7], u([[0, 0], [0, 0]])).tolist() == [[0, 0]]
    assert O.divide_and_round_q_last([13, 7], u([[1, 2], [1, 2]])).tolist() == [[0, 0]]
    assert O.divide_and_round_q_last([13, 7], u([[12, 11], [4, 3]])).tolist() == [[4, 3]]
    assert O.divide_and_round_q_last([13, 7], u([[6, 2], [5, 1]])).tolist() == [[3, 2]]
    q = [3, 5, 7, 11]
    assert (O.divide_and_round_q_last(q, u([[1, 2]] * 4)) == 0).all()
    out = O.divide_and_round_q_last(q, u([[0, 1], [0, 0], [4, 0], [5, 4]]))
    want = [(3, 2, 0), (5, 0, 1), (7, 5, 6)]
    for row, (p, w0, w1) in zip(out, want):  # the reference allows an error of one here
        assert (p + w0 - int(row[0])) % p <= 1 and (p + w1 - int(row[1])) % p <= 1


def test_scalar_kats():
    # native/tests/seal/util/uintarithsmallmod.cpp (MultiplyUIntMod :375-, BarrettReduce128 :142-) and
    # polyarithsmallmod.cpp DyadicProductCoeffMod :545-: the oracle's plain % arithmetic on the same operands
    lib = O.lib()
    import ctypes as C
    mods = O.coeff_modulus_create(64, [30])
    oc = O.Oracle(O.CKKS, 64, mods)
    x = np.zeros((2, 1, 64), dtype=np.uint64)
    y = np.zeros((2, 1, 64), dtype=np.uint64)
    x[0, 0, :3] = [1, 1, 1]
    x[1, 0, :3] = [2, 1, 2]
    y[0, 0, :3] = [2, 3, 4]
    y[1, 0, :3] = [2, 3, 4]
    m = oc.multiply(1, x, y)
    assert m[0, 0, :3].tolist() == [2, 3, 4] and m[2, 0, :3].tolist() == [4, 3, 8] and m[1, 0, :3].tolist() == [6, 6, 12]
