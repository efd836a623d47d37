"""Multi-chunk parity (-m gpu): the path bench.py times cuts a batch into key-switching chunks (op_relinearize /
op_multiply_relinearize / op_apply_galois, sb_engine.cu) and the *_host entry points cut it into staging chunks
(HostPipe, sb_api.cu).  These tests force >= 3 chunks with a ragged last chunk through sb200_context_set_limit and
compare EVERY ciphertext of the first, a middle and the last chunk with the reference itself (oracle/_ref =
Evaluator::multiply + relinearize_inplace / apply_galois, evaluator.cpp:2561-2867), and the remaining ciphertexts with
the single-chunk result of the same library."""
import numpy as np
import pytest

import oracle as O
import refseal as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsealref.so not present")


def sb():
    import seal_b200

    return seal_b200


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


def chunk_ranges(batch, chunk):
    return [(b0, min(b0 + chunk, batch)) for b0 in range(0, batch, chunk)]


def sampled(batch, chunk):
    """indices of every ciphertext of the first, a middle and the last chunk"""
    r = chunk_ranges(batch, chunk)
    assert len(r) >= 3 and (r[-1][1] - r[-1][0]) != chunk, "need >= 3 chunks with a ragged tail"
    pick = [r[0], r[len(r) // 2], r[-1]]
    return sorted({i for lo, hi in pick for i in range(lo, hi)})


def run_multichunk(n, bits, batch, chunk, rotate_step=None):
    import torch

    S = sb()
    mods = R.coeff_modulus_create(n, bits)
    L = len(mods) - 1
    rc = R.RefContext(R.CKKS, n, mods)
    ctx = S.Context(S.CKKS, n, mods)
    rk = ctx.load_key(rc.relin_key())
    a, b = device_rand(mods, n, (batch, 2), L, 11), device_rand(mods, n, (batch, 2), L, 12)
    idx = sampled(batch, chunk)

    # single-chunk results first (the shape every other parity test covers)
    one = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, one, L, batch)
    m3 = torch.empty((batch, 3, L, n), dtype=torch.int64, device="cuda")
    ctx.d_multiply(a, b, m3, L, batch)
    torch.cuda.synchronize()

    ctx.set_limit(ctx.LIMIT_KS_CHUNK, chunk)
    launches0 = ctx.launch_count
    out = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    launches_chunked = ctx.launch_count - launches0
    rel = torch.empty_like(a)
    ctx.d_relinearize(m3, rk, rel, L, batch)
    torch.cuda.synchronize()
    assert launches_chunked >= 3 * 5, "the chunked path did not run"
    assert torch.equal(out, one), "chunked multiply+relinearize differs from the single-chunk result"
    assert torch.equal(rel, one), "chunked relinearize(multiply) differs from the fused single-chunk result"
    want = {}
    for i in idx:
        want[i] = rc.multiply_relin(L, to_np(a[i]), to_np(b[i]))
        assert (to_np(out[i]) == want[i]).all(), f"multiply+relinearize, ciphertext {i} (chunk {i // chunk}) vs reference"

    if rotate_step is not None:
        e = rc.galois_elt_from_step(rotate_step)
        gk = ctx.load_key(rc.galois_key(e))
        rot = torch.empty_like(a)
        ctx.d_apply_galois(a, e, gk, rot, L, batch)
        ctx.set_limit(ctx.LIMIT_KS_CHUNK, 0)
        rot_one = torch.empty_like(a)
        ctx.d_apply_galois(a, e, gk, rot_one, L, batch)
        torch.cuda.synchronize()
        assert torch.equal(rot, rot_one)
        for i in idx:
            assert (to_np(rot[i]) == rc.apply_galois(L, to_np(a[i]), e)).all(), f"apply_galois, ciphertext {i} vs reference"

    # host pipeline (the e2e path of bench.py): >= 3 staging chunks, ragged tail; ciphertext words per op = 6 * L * n
    ctx.set_limit(ctx.LIMIT_KS_CHUNK, chunk)
    ctx.set_limit(ctx.LIMIT_HOST_STAGE_BYTES, chunk * 6 * L * n * 8)
    ha, hb = to_np(a), to_np(b)
    hout = ctx.multiply_relinearize(ha, hb, rk)
    assert (hout == to_np(one)).all(), "host pipeline (multi-chunk) differs from the device-resident result"
    ctx.set_limit(ctx.LIMIT_KS_CHUNK, 0)
    ctx.set_limit(ctx.LIMIT_HOST_STAGE_BYTES, chunk * 6 * L * n * 8)
    hout2 = ctx.multiply_relinearize(ha, hb, rk)
    assert (hout2 == hout).all()
    for i in idx:
        assert (hout[i] == want[i]).all()


@needs_ref
def test_multichunk_keyswitch_n8192_k4_vs_reference():
    # cfg2 shape: 11 ciphertexts in chunks of 3 -> 3,3,3,2
    run_multichunk(8192, [54, 54, 54, 54], batch=11, chunk=3, rotate_step=1)


@needs_ref
def test_multichunk_keyswitch_n65536_k32_vs_reference():
    # the headline shape (cfg5): 7 ciphertexts in chunks of 2 -> 2,2,2,1; 5 of them go through the CPU reference (~1.5 s each)
    run_multichunk(65536, [55] * 32, batch=7, chunk=2)


def test_multichunk_keyswitch_scratch_budget_vs_oracle():
    """the default chunking rule (scratch budget) with a small budget, checked against the oracle (no reference needed)"""
    import torch

    S = sb()
    n, bits, batch = 4096, [50, 50, 50, 50], 9
    mods = O.coeff_modulus_create(n, bits)
    k, L = len(mods), len(mods) - 1
    ctx = S.Context(S.CKKS, n, mods)
    oc = O.Oracle(O.CKKS, n, mods)
    key = to_np(device_rand(mods, n, (L, 2), k, 21))
    rk = ctx.load_key(key)
    a, b = device_rand(mods, n, (batch, 2), L, 22), device_rand(mods, n, (batch, 2), L, 23)
    one = torch.empty_like(a)
    ctx.d_multiply_relinearize(a, b, rk, one, L, batch)
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    ctx.d_multiply_relinearize(a, b, rk, one, L, batch)
    single = ctx.launch_count - l0
    ctx.set_limit(ctx.LIMIT_SCRATCH_BYTES, 4 << 20)  # a few ciphertexts per chunk
    out = torch.empty_like(a)
    l0 = ctx.launch_count
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    torch.cuda.synchronize()
    assert ctx.launch_count - l0 >= 3 * single, "expected at least three chunks"
    assert torch.equal(out, one)
    for i in (0, 4, 8):
        assert (to_np(out[i]) == oc.multiply_relin(L, to_np(a[i]), to_np(b[i]), key)).all()


def test_selftest_rates_and_profile_work_counters():
    """the in-process ceilings (sb200_selftest_rate) and the work counters of the profile (butterflies, multiply-accumulates per
    launch: the second ceiling of SURVEY 8d) are sane"""
    import torch

    S = sb()
    n, bits, batch = 8192, [54, 54, 54, 55], 5
    mods = O.coeff_modulus_create(n, bits)
    k, L = len(mods), len(mods) - 1
    key = to_np(device_rand(mods, n, (L, 2), k, 31))
    a, b = device_rand(mods, n, (batch, 2), L, 32), device_rand(mods, n, (batch, 2), L, 33)
    ctx = S.Context(S.CKKS, n, mods)
    rk = ctx.load_key(key)
    out = torch.empty_like(a)
    # (a) the 64-bit digit-transform path
    ctx.set_limit(ctx.LIMIT_KS_ALGORITHM, 0)
    ctx.profile(True)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    torch.cuda.synchronize()
    work = {r[0]: r for r in ctx.profile_read_work()}
    ctx.profile(False)
    assert work["ks_local_mac"][4] == 0.5 * batch * L * L * n * 8 and work["ks_local_mac"][5] == 2.0 * batch * (L + 1) * L * n
    assert work["ks_digit_ntt:col"][4] == 0.5 * batch * L * L * n * (13 - 8)
    rates = [ctx.selftest_rate(kind) for kind in range(4)]
    assert all(r > 1e9 for r in rates), rates
    # (b) the integer path (default): 32-bit butterflies and multiply-accumulates are counted separately
    ctx.set_limit(ctx.LIMIT_KS_ALGORITHM, 2)
    S5 = len(ctx.ksint_primes())
    out64 = out.clone()
    out.zero_()
    ctx.profile(True)
    ctx.d_multiply_relinearize(a, b, rk, out, L, batch)
    torch.cuda.synchronize()
    work = {r[0]: r for r in ctx.profile_read_work32()}
    ctx.profile(False)
    assert torch.equal(out, out64)
    assert work["ks32_mac"][7] == 2.0 * batch * (L + 1) * L * S5 * n and work["ks32_mac"][5] == 0
    assert work["ks32_fwd_local"][6] == 0.5 * batch * L * S5 * n * 12
    assert work["ks32_inv_local"][6] == 0.5 * batch * 2 * (L + 1) * S5 * n * 12
    rates = [ctx.selftest_rate(kind) for kind in (10, 11, 12, 13, 14, 15)]
    assert all(r > 1e9 for r in rates), rates
    oc = O.Oracle(O.CKKS, n, mods)
    assert (to_np(out[batch - 1]) == oc.multiply_relin(L, to_np(a[batch - 1]), to_np(b[batch - 1]), key)).all()
