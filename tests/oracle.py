"""ctypes driver for oracle/liboracle.so (the plain-C restatement, oracle/seal_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "..", "oracle")
_LIB_PATH = os.path.join(_DIR, "liboracle.so")
BFV, CKKS, BGV = 1, 2, 3
_u64p = C.POINTER(C.c_uint64)
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_DIR, "seal_oracle.c")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_blake2xb_stream.argtypes = [_u64p, C.c_size_t, _u64p]
        L.orc_expand_seed.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        L.orc_ckks_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double, _u64p]
        L.orc_ckks_decode.argtypes = [C.c_void_p, C.c_size_t, _u64p, C.c_double, C.c_void_p]
        L.orc_encrypt_zero_asymmetric.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p]
        L.orc_encrypt_zero_symmetric.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, C.c_int, _u64p]
        L.orc_create.argtypes = [C.c_int, C.c_size_t, _u64p, C.c_size_t, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_get_primes.argtypes = [C.c_uint64, C.c_int, C.c_size_t, _u64p]
        L.orc_minimal_primitive_root.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.orc_coeff_modulus_create.argtypes = [C.c_size_t, C.POINTER(C.c_int), C.c_size_t, _u64p]
        L.orc_ntt_tables.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p, _u64p]
        L.orc_base_bsk.restype = C.c_size_t
        L.orc_base_bsk.argtypes = [C.c_void_p, C.c_size_t, _u64p]
        L.orc_ntt_row.argtypes = [C.c_void_p, C.c_size_t, _u64p]
        L.orc_intt_row.argtypes = [C.c_void_p, C.c_size_t, _u64p]
        L.orc_ntt_forward.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _u64p]
        L.orc_ntt_inverse.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _u64p]
        L.orc_ckks_multiply.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p]
        L.orc_bfv_multiply.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p]
        L.orc_switch_key.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p]
        L.orc_relinearize.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p, _u64p]
        L.orc_rescale.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        L.orc_bfv_mod_switch.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        L.orc_bgv_mod_switch.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        L.orc_apply_galois.argtypes = [C.c_void_p, C.c_size_t, _u64p, C.c_uint32, _u64p, _u64p]
        L.orc_galois_elt_from_step.restype = C.c_uint32
        L.orc_galois_elt_from_step.argtypes = [C.c_size_t, C.c_int]
        L.orc_galois_coeff_row.argtypes = [C.c_size_t, C.c_uint64, C.c_uint32, _u64p, _u64p]
        L.orc_galois_ntt_row.argtypes = [C.c_size_t, C.c_uint32, _u64p, _u64p]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def get_primes(factor, bits, count):
    out = np.zeros(count, dtype=np.uint64)
    if lib().orc_get_primes(factor, bits, count, _p(out)):
        raise RuntimeError("not enough primes")
    return [int(x) for x in out]


def coeff_modulus_create(n, bits):
    out = np.zeros(len(bits), dtype=np.uint64)
    b = (C.c_int * len(bits))(*bits)
    if lib().orc_coeff_modulus_create(n, b, len(bits), _p(out)):
        raise RuntimeError("coeff_modulus_create failed")
    return [int(x) for x in out]


def minimal_primitive_root(degree, q):
    r = C.c_uint64(0)
    if lib().orc_minimal_primitive_root(degree, q, C.byref(r)):
        raise RuntimeError("no primitive root")
    return r.value


def galois_elt_from_step(n, step):
    return int(lib().orc_galois_elt_from_step(n, step))


def galois_coeff_row(n, q, elt, row):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_galois_coeff_row(n, q, elt, _p(np.ascontiguousarray(row)), _p(out))
    return out


def galois_ntt_row(n, elt, row):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_galois_ntt_row(n, elt, _p(np.ascontiguousarray(row)), _p(out))
    return out


class Oracle:
    def __init__(self, scheme, n, moduli, plain_modulus=0):
        self.scheme, self.n, self.moduli, self.k, self.t = scheme, n, list(moduli), len(moduli), plain_modulus
        m = np.array(self.moduli, dtype=np.uint64)
        self.h = lib().orc_create(scheme, n, _p(m), self.k, plain_modulus)
        if not self.h:
            raise RuntimeError("orc_create failed")

    def ckks_encode(self, L, values, scale):
        """CKKSEncoder::encode of a complex vector -> [L][n] NTT form, or None for the reference's invalid_argument cases"""
        v = np.ascontiguousarray(values, dtype=np.complex128)
        out = np.zeros((L, self.n), dtype=np.uint64)
        return None if lib().orc_ckks_encode(self.h, L, v.ctypes.data, v.size, float(scale), _p(out)) else out

    def ckks_decode(self, L, plain, scale):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros(self.n // 2, dtype=np.complex128)
        return None if lib().orc_ckks_decode(self.h, L, _p(plain), float(scale), out.ctypes.data) else out

    def encrypt_zero_asymmetric(self, pk, seed, L=None):
        """public-key encryption of zero with the PRNG seeded by `seed` (8 words): [2][L][n]"""
        pk = np.ascontiguousarray(pk, dtype=np.uint64)
        seed = np.ascontiguousarray(seed, dtype=np.uint64)
        L = L or (self.k - 1 if self.k > 1 else 1)
        out = np.zeros((2, L, self.n), dtype=np.uint64)
        lib().orc_encrypt_zero_asymmetric(self.h, L, _p(pk), _p(seed), _p(out))
        return out

    def encrypt_zero_symmetric(self, sk, seed, save_seed, L=None):
        """encrypt_zero_symmetric with the bootstrap PRNG seeded by `seed` (8 words): [2][L][n] (default: the first data level)"""
        sk = np.ascontiguousarray(sk, dtype=np.uint64)
        seed = np.ascontiguousarray(seed, dtype=np.uint64)
        L = L or (self.k - 1 if self.k > 1 else 1)
        out = np.zeros((2, L, self.n), dtype=np.uint64)
        lib().orc_encrypt_zero_symmetric(self.h, L, _p(sk), _p(seed), int(save_seed), _p(out))
        return out

    @staticmethod
    def blake2xb_stream(seed, words):
        """the first `words` 64-bit words of Blake2xbPRNG(seed)"""
        seed = np.ascontiguousarray(seed, dtype=np.uint64)
        out = np.zeros(words, dtype=np.uint64)
        lib().orc_blake2xb_stream(_p(seed), words, _p(out))
        return out

    def expand_seed(self, L, seed):
        """Ciphertext::expand_seed: the polynomial [L][n] a 64-byte PRNG seed (8 words) expands into (Blake2xbPRNG)"""
        seed = np.ascontiguousarray(seed, dtype=np.uint64)
        assert seed.shape == (8,)
        out = np.zeros((L, self.n), dtype=np.uint64)
        lib().orc_expand_seed(self.h, L, _p(seed), _p(out))
        return out

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ntt_tables(self, i):
        n = self.n
        rp, irp = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        root, inv_n = C.c_uint64(0), C.c_uint64(0)
        assert lib().orc_ntt_tables(self.h, i, C.byref(root), _p(rp), _p(irp), C.byref(inv_n)) == 0
        return root.value, rp, irp, inv_n.value

    def base_bsk(self, L):
        out = np.zeros(self.k + 4, dtype=np.uint64)
        cnt = lib().orc_base_bsk(self.h, L, _p(out))
        return [int(x) for x in out[:cnt]]

    def ntt_row(self, i, row):
        r = np.ascontiguousarray(row).copy()
        lib().orc_ntt_row(self.h, i, _p(r))
        return r

    def intt_row(self, i, row):
        r = np.ascontiguousarray(row).copy()
        lib().orc_intt_row(self.h, i, _p(r))
        return r

    def ntt_forward(self, L, data):
        d = np.ascontiguousarray(data).copy()
        lib().orc_ntt_forward(self.h, L, d.shape[0], _p(d))
        return d

    def ntt_inverse(self, L, data):
        d = np.ascontiguousarray(data).copy()
        lib().orc_ntt_inverse(self.h, L, d.shape[0], _p(d))
        return d

    def multiply(self, L, a, b):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        sa, sb = a.shape[0], b.shape[0]
        out = np.zeros((sa + sb - 1, L, self.n), dtype=np.uint64)
        sized = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p, _u64p, _u64p]
        lib().orc_ckks_multiply_sized.argtypes = lib().orc_bfv_multiply_sized.argtypes = sized
        if self.scheme != BFV:  # CKKS and BGV: the NTT-form tensor
            lib().orc_ckks_multiply_sized(self.h, L, sa, sb, _p(a), _p(b), _p(out))
        else:
            assert lib().orc_bfv_multiply_sized(self.h, L, sa, sb, _p(a), _p(b), _p(out)) == 0
        return out

    def linear(self, mode, L, a, b=None):
        lib().orc_linear.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, _u64p, _u64p, _u64p]
        a = np.ascontiguousarray(a)
        out = np.zeros_like(a)
        bb = np.ascontiguousarray(b) if b is not None else a
        lib().orc_linear(self.h, mode, L, a.shape[0], _p(a), _p(bb), _p(out))
        return out

    def multiply_plain(self, L, a, plain):
        lib().orc_multiply_plain_ntt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _u64p, _u64p, _u64p]
        a, plain = np.ascontiguousarray(a), np.ascontiguousarray(plain)
        out = np.zeros_like(a)
        lib().orc_multiply_plain_ntt(self.h, L, a.shape[0], _p(a), _p(plain), _p(out))
        return out

    def decrypt(self, L, ct, sk, correction_factor=1):
        """Decryptor::decrypt with the secret key sk [k][n] (NTT form): n words (BFV / BGV) or [L][n] (CKKS)"""
        ct, sk = np.ascontiguousarray(ct), np.ascontiguousarray(sk)
        size = ct.shape[0]
        if self.scheme == CKKS:
            lib().orc_decrypt_phase.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, _u64p, _u64p, _u64p]
            out = np.zeros((L, self.n), dtype=np.uint64)
            lib().orc_decrypt_phase(self.h, L, size, 1, _p(ct), _p(sk), _p(out))
            return out
        out = np.zeros(self.n, dtype=np.uint64)
        if self.scheme == BFV:
            lib().orc_bfv_decrypt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _u64p, _u64p, _u64p]
            assert lib().orc_bfv_decrypt(self.h, L, size, _p(ct), _p(sk), _p(out)) == 0
        else:
            lib().orc_bgv_decrypt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, _u64p, _u64p, _u64p]
            assert lib().orc_bgv_decrypt(self.h, L, size, correction_factor, _p(ct), _p(sk), _p(out)) == 0
        return out

    def batch_codec(self, data, decode):
        lib().orc_batch_codec.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p]
        data = np.ascontiguousarray(data, dtype=np.uint64)
        out = np.zeros(self.n, dtype=np.uint64)
        assert lib().orc_batch_codec(self.h, int(decode), _p(data), _p(out)) == 0
        return out

    def plain_to_ntt(self, L, plain):
        lib().orc_plain_to_ntt.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros((L, self.n), dtype=np.uint64)
        lib().orc_plain_to_ntt(self.h, L, _p(plain), _p(out))
        return out

    def multiply_plain_coeff(self, L, a, plain, ct_is_ntt):
        lib().orc_multiply_plain_coeff.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, _u64p, _u64p, _u64p]
        a, plain = np.ascontiguousarray(a), np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros_like(a)
        lib().orc_multiply_plain_coeff(self.h, L, a.shape[0], int(ct_is_ntt), _p(a), _p(plain), _p(out))
        return out

    def add_plain_coeff(self, L, a, plain, subtract=False, correction_factor=1):
        lib().orc_add_plain_coeff.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_uint64, _u64p, _u64p, _u64p]
        a, plain = np.ascontiguousarray(a), np.ascontiguousarray(plain, dtype=np.uint64)
        out = np.zeros_like(a)
        lib().orc_add_plain_coeff(self.h, L, a.shape[0], int(subtract), correction_factor, _p(a), _p(plain), _p(out))
        return out

    def relinearize(self, L, c3, key):
        out = np.zeros((2, L, self.n), dtype=np.uint64)
        lib().orc_relinearize(self.h, L, _p(c3), _p(key), _p(out))
        return out

    def multiply_relin(self, L, a, b, key):
        return self.relinearize(L, self.multiply(L, a, b), key)

    def rescale(self, L, c2):
        out = np.zeros((2, L - 1, self.n), dtype=np.uint64)
        lib().orc_rescale(self.h, L, _p(c2), _p(out))
        return out

    def bfv_mod_switch(self, L, c2):
        out = np.zeros((2, L - 1, self.n), dtype=np.uint64)
        lib().orc_bfv_mod_switch(self.h, L, _p(c2), _p(out))
        return out

    def bgv_mod_switch(self, L, c2):
        out = np.zeros((2, L - 1, self.n), dtype=np.uint64)
        lib().orc_bgv_mod_switch(self.h, L, _p(c2), _p(out))
        return out

    def apply_galois(self, L, c2, elt, key):
        out = np.zeros((2, L, self.n), dtype=np.uint64)
        lib().orc_apply_galois(self.h, L, _p(c2), elt, _p(key), _p(out))
        return out


# ---- standalone RNS building blocks (explicit bases) ----
def _setup_rns():
    L = lib()
    L.orc_fastbconv_array.argtypes = [_u64p, C.c_size_t, _u64p, C.c_size_t, _u64p, C.c_size_t, _u64p]
    L.orc_behz_base.restype = C.c_size_t
    L.orc_behz_base.argtypes = [C.c_size_t, _u64p, C.c_size_t, C.c_uint64, _u64p]
    for f in (L.orc_behz_fastbconv_m_tilde, L.orc_behz_sm_mrq, L.orc_behz_fast_floor, L.orc_behz_fastbconv_sk):
        f.argtypes = [C.c_size_t, _u64p, C.c_size_t, C.c_uint64, _u64p, _u64p]
    L.orc_divide_and_round_q_last.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p]
    return L


def fastbconv_array(ibase, obase, x):
    """BaseConverter::fast_convert_array: x [ni][n] -> [no][n]"""
    L = _setup_rns()
    x = np.ascontiguousarray(x, dtype=np.uint64)
    n = x.shape[1]
    out = np.zeros((len(obase), n), dtype=np.uint64)
    L.orc_fastbconv_array(_p(np.array(ibase, dtype=np.uint64)), len(ibase), _p(np.array(obase, dtype=np.uint64)), len(obase), _p(x), n, _p(out))
    return out


def behz_base(n, q, t=0):
    L = _setup_rns()
    out = np.zeros(len(q) + 4, dtype=np.uint64)
    cnt = L.orc_behz_base(n, _p(np.array(q, dtype=np.uint64)), len(q), t, _p(out))
    return [int(v) for v in out[:cnt]]


def _behz(fn, n, q, t, x, rows_out):
    L = _setup_rns()
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.zeros((rows_out, n), dtype=np.uint64)
    assert getattr(L, fn)(n, _p(np.array(q, dtype=np.uint64)), len(q), t, _p(x), _p(out)) == 0
    return out


def behz_fastbconv_m_tilde(n, q, x, t=0):
    return _behz("orc_behz_fastbconv_m_tilde", n, q, t, x, len(behz_base(n, q, t)) + 1)


def behz_sm_mrq(n, q, x, t=0):
    return _behz("orc_behz_sm_mrq", n, q, t, x, len(behz_base(n, q, t)))


def behz_fast_floor(n, q, x, t=0):
    return _behz("orc_behz_fast_floor", n, q, t, x, len(behz_base(n, q, t)))


def behz_fastbconv_sk(n, q, x, t=0):
    return _behz("orc_behz_fastbconv_sk", n, q, t, x, len(q))


def divide_and_round_q_last(q, x):
    L = _setup_rns()
    x = np.ascontiguousarray(x, dtype=np.uint64).copy()
    L.orc_divide_and_round_q_last(_p(np.array(q, dtype=np.uint64)), len(q), x.shape[1], _p(x))
    return x[:-1]
