"""BASELINE.json configs at their named shapes (-m gpu).  Full batches run on the GPU; the reference (oracle/_ref, where
present) checks sampled ciphertexts word for word, and size-independent properties cover the rest of the batch:
transform round trips, commutativity of multiply, fused == unfused, batch == singles."""
import numpy as np
import pytest

import oracle as O
import refseal as R
from common import rand_ct

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsealref.so not present")


def sb():
    import seal_b200

    return seal_b200


def torch_slab(a):
    import torch

    return torch.from_numpy(a.view(np.int64)).cuda()


def to_np(t):
    return t.cpu().numpy().view(np.uint64)


def device_rand(mods, n, shape_prefix, L, seed):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.empty((*shape_prefix, L, n), dtype=torch.int64, device="cuda")
    for i in range(L):
        t[..., i, :] = torch.randint(0, mods[i], (*shape_prefix, n), generator=g, dtype=torch.int64, device="cuda")
    return t


@needs_ref
def test_cfg2_ckks_n8192_k4_batch1024():
    import torch

    n, bits, batch = 8192, [54, 54, 54, 54], 1024
    mods = R.coeff_modulus_create(n, bits)
    rc = R.RefContext(R.CKKS, n, mods)
    ctx = sb().Context(sb().CKKS, n, mods)
    rk = ctx.load_key(rc.relin_key())
    L = 3
    a, b = device_rand(mods, n, (batch, 2), L, 1), device_rand(mods, n, (batch, 2), L, 2)
    out = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    out_ba = torch.empty_like(a)
    ctx.d_multiply_relinearize(b, a, rk, out_ba, L, batch)
    m3 = torch.empty((batch, 3, L, n), dtype=torch.int64, device="cuda")
    ctx.d_multiply(a, b, m3, L, batch)
    out_unfused = torch.empty_like(a)
    ctx.d_relinearize(m3, rk, out_unfused, L, batch)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ba), "multiply is commutative, word for word"
    assert torch.equal(out, out_unfused), "fused multiply+relinearize == relinearize(multiply)"
    for i in (0, 1, 511, 1023):
        want = rc.multiply_relin(L, to_np(a[i]), to_np(b[i]))
        assert (to_np(out[i]) == want).all()
        # batch == single
        assert (ctx.multiply_relinearize(to_np(a[i]), to_np(b[i]), rk) == want).all()
    # NTT round trip over the whole batch
    x = a.clone()
    ctx.d_ntt_inverse(x, L, 2, batch)
    ctx.d_ntt_forward(x, L, 2, batch)
    torch.cuda.synchronize()
    assert torch.equal(x, a)


@needs_ref
def test_cfg3_ckks_n32768_k16_chain_depth8():
    import torch

    n, batch, depth = 32768, 16, 8
    mods = R.coeff_modulus_bfv_default(n)  # 15 x 55-bit + 56-bit (util/globals.cpp:66-72)
    assert len(mods) == 16
    rc = R.RefContext(R.CKKS, n, mods)
    ctx = sb().Context(sb().CKKS, n, mods)
    rk = ctx.load_key(rc.relin_key())
    L = 15
    a, b = device_rand(mods, n, (batch, 2), L, 3), device_rand(mods, n, (batch, 2), L, 4)
    ra, rb = [to_np(a[i]) for i in (0, batch - 1)], [to_np(b[i]) for i in (0, batch - 1)]
    for _ in range(depth):  # a <- rescale(relin(a*b)); b <- mod_switch_to_next(b)   
        prod = torch.empty((batch, 2, L, n), dtype=torch.int64, device="cuda")
        ctx.d_multiply_relinearize(a, b, rk, prod, L, batch)
        a = torch.empty((batch, 2, L - 1, n), dtype=torch.int64, device="cuda")
        ctx.d_rescale_to_next(prod, a, L, batch)
        nb = torch.empty((batch, 2, L - 1, n), dtype=torch.int64, device="cuda")
        ctx.d_mod_switch_to_next(b, nb, L, batch)
        b = nb
        for s in range(2):
            ra[s] = rc.rescale(L, rc.multiply_relin(L, ra[s], rb[s]))
            rb[s] = rc.mod_switch(L, rb[s])
        L -= 1
    torch.cuda.synchronize()
    assert L == 7
    for s, i in enumerate((0, batch - 1)):
        assert (to_np(a[i]) == ra[s]).all()
        assert (to_np(b[i]) == rb[s]).all()


@needs_ref
def test_cfg4_bfv_n16384_k8_rotate_rows_sweep():
    import torch

    n, batch = 16384, 8
    mods = R.coeff_modulus_create(n, [54] * 8)
    t = R.plain_modulus_batching(n, 20)
    rb = R.RefContext(R.BFV, n, mods, t)
    ctx = sb().Context(sb().BFV, n, mods, t)
    L = 7
    a = device_rand(mods, n, (batch, 2), L, 5)
    a0 = to_np(a[0])
    # all Galois elements of create_galois_keys(): m-1 and 3^(+-2^k) (util/galois.cpp:106-131) = steps 0, +-2^k
    steps = [0] + [s * (1 << k) for k in range(13) for s in (1, -1) if (1 << k) < n // 2]
    elts = sorted({rb.galois_elt_from_step(s) for s in steps})
    assert len(elts) == 26  # get_elts_all lists 27 entries; 3^(2^12) = 3^-(2^12) mod 2n coincide
    out = torch.empty_like(a)
    for e in elts:
        gk = ctx.load_key(rb.galois_key(e))
        ctx.d_apply_galois(a, e, gk, out, L, batch)
        torch.cuda.synchronize()
        assert (to_np(out[0]) == rb.apply_galois(L, a0, e)).all(), f"galois element {e}"


@needs_ref
def test_cfg5_ckks_n65536_k32_sample():
    import torch

    n, batch = 65536, 6
    mods = R.coeff_modulus_create(n, [55] * 32)
    assert mods == sb().coeff_modulus_create(n, [55] * 32)
    rc = R.RefContext(R.CKKS, n, mods)
    ctx = sb().Context(sb().CKKS, n, mods)
    rk = ctx.load_key(rc.relin_key())
    L = 31
    a, b = device_rand(mods, n, (batch, 2), L, 6), device_rand(mods, n, (batch, 2), L, 7)
    out = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    out_ba = torch.empty_like(a)
    ctx.d_multiply_relinearize(b, a, rk, out_ba, L, batch)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ba)
    i = batch - 1
    assert (to_np(out[i]) == rc.multiply_relin(L, to_np(a[i]), to_np(b[i]))).all()
    # rotate one step + rescale on the same sample
    e = rc.galois_elt_from_step(1)
    gk = ctx.load_key(rc.galois_key(e))
    rot = ctx.apply_galois(to_np(a[i]), e, gk)
    assert (rot == rc.apply_galois(L, to_np(a[i]), e)).all()
    assert (ctx.rescale_to_next(to_np(a[i])) == rc.rescale(L, to_np(a[i]))).all()


def test_full_size_properties_without_reference():
    # runs even where the reference library is absent: oracle on one sampled ciphertext + structural properties
    import torch

    n, bits, batch = 16384, [50, 50, 50, 50, 50], 64
    mods = O.coeff_modulus_create(n, bits)
    ctx = sb().Context(sb().CKKS, n, mods)
    L, k = 4, 5
    key = to_np(device_rand(mods, n, (L, 2), k, 8))
    rk = ctx.load_key(key)
    a, b = device_rand(mods, n, (batch, 2), L, 9), device_rand(mods, n, (batch, 2), L, 10)
    out = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    out_ba = torch.empty_like(a)
    ctx.d_multiply_relinearize(b, a, rk, out_ba, L, batch)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ba)
    oc = O.Oracle(O.CKKS, n, mods)
    assert (to_np(out[17]) == oc.multiply_relin(L, to_np(a[17]), to_np(b[17]), key)).all()
    res = torch.empty((batch, 2, L - 1, n), dtype=torch.int64, device="cuda")
    ctx.d_rescale_to_next(out, res, L, batch)
    torch.cuda.synchronize()
    assert (to_np(res[17]) == oc.rescale(L, to_np(out[17]))).all()
