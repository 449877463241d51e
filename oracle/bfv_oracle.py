"""ctypes binding of the CPU oracle (oracle/liboracle_bfv.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (abc_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_bfv.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build():
    """Compile the oracle if a compiler is here (the GPU box uses the prebuilt .so)."""
    src = os.path.join(_HERE, "bfv_oracle.c")
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "bfv_oracle.h"))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    build()
    lib = C.CDLL(_SO)
    vp, sz, u64, i32, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint32
    sig = {
        "obfv_create": (vp, [sz, vp, sz, u64]),
        "obfv_create_aux": (vp, [sz, vp, sz, u64, i32, sz]),
        "obfv_destroy": (None, [vp]),
        "obfv_N": (sz, [vp]), "obfv_k": (sz, [vp]), "obfv_L": (sz, [vp]), "obfv_t": (u64, [vp]),
        "obfv_primes": (None, [vp, u64p]),
        "obfv_nbsk": (sz, [vp]),
        "obfv_aux_primes": (None, [vp, C.POINTER(u64), C.POINTER(u64), u64p]),
        "obfv_psi": (u64, [vp, sz]), "obfv_psi_t": (u64, [vp]),
        "obfv_get_primes": (sz, [sz, i32, sz, u64p]),
        "obfv_ntt_fwd": (None, [vp, sz, u64p]), "obfv_ntt_inv": (None, [vp, sz, u64p]),
        "obfv_keygen": (None, [vp, u64]),
        "obfv_keygen_select": (None, [vp, u64, u32p, sz]),
        "obfv_secret_key": (vp, [vp]), "obfv_public_key": (vp, [vp]), "obfv_relin_key": (vp, [vp]),
        "obfv_galois_key": (vp, [vp, u32]),
        "obfv_galois_elts": (sz, [vp, u32p, sz]),
        "obfv_elt_from_step": (u32, [vp, i32]),
        "obfv_encode": (None, [vp, i64p, u64p]), "obfv_decode": (None, [vp, u64p, i64p]),
        "obfv_encrypt": (None, [vp, u64p, u64, u64p]),
        "obfv_decrypt": (None, [vp, u64p, sz, u64p]),
        "obfv_noise_budget": (C.c_int, [vp, u64p, sz]),
        "obfv_add": (None, [vp, u64p, u64p, u64p]), "obfv_sub": (None, [vp, u64p, u64p, u64p]),
        "obfv_negate": (None, [vp, u64p, u64p]),
        "obfv_add_plain": (None, [vp, u64p, u64p, u64p]), "obfv_sub_plain": (None, [vp, u64p, u64p, u64p]),
        "obfv_multiply_plain": (None, [vp, u64p, u64p, u64p]),
        "obfv_multiply": (None, [vp, u64p, u64p, u64p]),
        "obfv_relinearize": (None, [vp, u64p, u64p]),
        "obfv_apply_galois": (None, [vp, u64p, u32, u64p]),
        "obfv_rotate_rows": (i32, [vp, u64p, i32, u64p]),
        "obfv_rotate_keyswitch_count": (i32, [vp, i32]),
        "obfv_behz_lift": (None, [vp, u64p, u64p]),
        "obfv_behz_scale": (None, [vp, u64p, u64p, u64p]),
        "obfv_switch_key": (None, [vp, u64p, u64p, vp]),
        "obfv_rng": (u64, [u64, vp, u64, u64, u64, u64]),
        "obfv_set_rng_key": (None, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def get_primes(N, bits, count):
    out = np.zeros(count, dtype=np.uint64)
    n = lib().obfv_get_primes(N, bits, count, out)
    return [int(v) for v in out[:n]]


class Oracle:
    """SEAL-3.6.5-restatement BFV context.  Mirrors what SealCiphertextFactory sets up
    (/root/reference/src/runtime/SealCiphertextFactory.cpp:72-100)."""

    def __init__(self, N, primes=None, t=0, seed=None, galois_steps=None, aux=None, rng_key=None):
        """aux = (bits, count): test hook, BEHZ auxiliary base of `count` + 1 primes of `bits` bits instead of SEAL's
        61-bit ones (obfv_create_aux); None = SEAL's base.  rng_key: the sampler's 32-byte ChaCha20 key itself (mirror of
        abc_set_rng_key); otherwise the key is expanded from `seed`."""
        L_ = lib()
        arr = None if primes is None else np.asarray(primes, dtype=np.uint64)
        ptr, n = (None, 0) if arr is None else (arr.ctypes.data, len(arr))
        if aux is None:
            self._c = L_.obfv_create(N, ptr, n, t)
        else:
            self._c = L_.obfv_create_aux(N, ptr, n, t, int(aux[0]), int(aux[1]))
        if not self._c:
            raise ValueError("invalid BFV parameters")
        self.N, self.k, self.L, self.t = N, L_.obfv_k(self._c), L_.obfv_L(self._c), L_.obfv_t(self._c)
        self.nbsk = L_.obfv_nbsk(self._c)
        q = np.zeros(self.k, dtype=np.uint64)
        L_.obfv_primes(self._c, q)
        self.primes = [int(v) for v in q]
        self.ct_words = 2 * self.L * N
        self.seed = None
        if rng_key is not None:
            if len(rng_key) != 32:
                raise ValueError("rng_key must be 32 bytes")
            self._rng_key = (C.c_ubyte * 32)(*rng_key)
            L_.obfv_set_rng_key(self._c, self._rng_key)
            if seed is None:
                seed = 0
        if seed is not None:
            self.keygen(seed, galois_steps)

    def __del__(self):
        if getattr(self, "_c", None):
            lib().obfv_destroy(self._c)
            self._c = None

    # -- parameters
    def aux_primes(self):
        msk, g = C.c_uint64(), C.c_uint64()
        B = np.zeros(self.nbsk - 1, dtype=np.uint64)
        lib().obfv_aux_primes(self._c, C.byref(msk), C.byref(g), B)
        return msk.value, g.value, [int(b) for b in B]

    def psi(self, i):
        return lib().obfv_psi(self._c, i)

    def psi_t(self):
        return lib().obfv_psi_t(self._c)

    def ntt_fwd(self, idx, limb):
        out = np.ascontiguousarray(limb, dtype=np.uint64).copy()
        lib().obfv_ntt_fwd(self._c, idx & (2**64 - 1), out)
        return out

    def ntt_inv(self, idx, limb):
        out = np.ascontiguousarray(limb, dtype=np.uint64).copy()
        lib().obfv_ntt_inv(self._c, idx & (2**64 - 1), out)
        return out

    # -- keys
    def keygen(self, seed, galois_steps=None):
        self.seed = seed
        if galois_steps is None:
            lib().obfv_keygen(self._c, seed)
        else:
            elts = np.asarray([self.elt_from_step(s) for s in galois_steps], dtype=np.uint32)
            lib().obfv_keygen_select(self._c, seed, elts, elts.size)

    def _view(self, ptr, shape):
        n = int(np.prod(shape))
        buf = (C.c_uint64 * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint64).reshape(shape)

    def secret_key(self):
        return self._view(lib().obfv_secret_key(self._c), (self.k, self.N))

    def public_key(self):
        return self._view(lib().obfv_public_key(self._c), (2, self.k, self.N))

    def relin_key(self):
        return self._view(lib().obfv_relin_key(self._c), (self.L, 2, self.k, self.N))

    def galois_key(self, elt):
        p = lib().obfv_galois_key(self._c, elt)
        return None if not p else self._view(p, (self.L, 2, self.k, self.N))

    def galois_elts(self):
        out = np.zeros(64, dtype=np.uint32)
        n = lib().obfv_galois_elts(self._c, out, 64)
        return [int(v) for v in out[:n]]

    def elt_from_step(self, step):
        return lib().obfv_elt_from_step(self._c, step)

    # -- encode / decode (BatchEncoder)
    def expand(self, values):
        """SealCiphertextFactory::expandVector: pad to N slots with the LAST value."""
        v = list(values)
        if len(v) > self.N:
            raise RuntimeError("Cannot encode %d elements in a ciphertext of size %d." % (len(v), self.N))
        return np.asarray(v + [v[-1]] * (self.N - len(v)), dtype=np.int64)

    def encode(self, slots):
        plain = np.zeros(self.N, dtype=np.uint64)
        lib().obfv_encode(self._c, np.ascontiguousarray(slots, dtype=np.int64), plain)
        return plain

    def decode(self, plain):
        slots = np.zeros(self.N, dtype=np.int64)
        lib().obfv_decode(self._c, np.ascontiguousarray(plain, dtype=np.uint64), slots)
        return slots

    # -- ciphertext ops; ciphertexts are uint64 arrays [size][L][N]
    def _ct(self, size=2):
        return np.zeros((size, self.L, self.N), dtype=np.uint64)

    def encrypt(self, plain, nonce):
        ct = self._ct()
        lib().obfv_encrypt(self._c, np.ascontiguousarray(plain, dtype=np.uint64), nonce, ct)
        return ct

    def encrypt_slots(self, values, nonce):
        return self.encrypt(self.encode(self.expand(values)), nonce)

    def decrypt(self, ct):
        plain = np.zeros(self.N, dtype=np.uint64)
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        lib().obfv_decrypt(self._c, ct, ct.shape[0], plain)
        return plain

    def noise_budget(self, ct):
        """Decryptor::invariant_noise_budget, SealCiphertext::noiseBits (SealCiphertext.cpp:80-83)."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        return int(lib().obfv_noise_budget(self._c, ct, ct.shape[0]))

    def decrypt_slots(self, ct):
        return self.decode(self.decrypt(ct))

    def _bin(self, fn, a, b):
        out = self._ct()
        fn(self._c, np.ascontiguousarray(a), np.ascontiguousarray(b), out)
        return out

    def add(self, a, b):
        return self._bin(lib().obfv_add, a, b)

    def sub(self, a, b):
        return self._bin(lib().obfv_sub, a, b)

    def negate(self, a):
        out = self._ct()
        lib().obfv_negate(self._c, np.ascontiguousarray(a), out)
        return out

    def add_plain(self, a, plain):
        return self._bin(lib().obfv_add_plain, a, plain)

    def sub_plain(self, a, plain):
        return self._bin(lib().obfv_sub_plain, a, plain)

    def multiply_plain(self, a, plain):
        return self._bin(lib().obfv_multiply_plain, a, plain)

    def multiply(self, a, b):
        out = self._ct(3)
        lib().obfv_multiply(self._c, np.ascontiguousarray(a), np.ascontiguousarray(b), out)
        return out

    def relinearize(self, ct3):
        out = self._ct()
        lib().obfv_relinearize(self._c, np.ascontiguousarray(ct3), out)
        return out

    def mul_relin(self, a, b):
        """SealCiphertext::multiply = multiply then relinearize_inplace
        (/root/reference/src/runtime/SealCiphertext.cpp:102-107)."""
        return self.relinearize(self.multiply(a, b))

    def apply_galois(self, a, elt):
        out = self._ct()
        lib().obfv_apply_galois(self._c, np.ascontiguousarray(a), elt, out)
        return out

    def rotate_rows(self, a, steps):
        out = self._ct()
        r = lib().obfv_rotate_rows(self._c, np.ascontiguousarray(a), steps, out)
        if r:
            raise ValueError("rotate_rows: invalid step count %d (code %d)" % (steps, r))
        return out

    def rotate_keyswitch_count(self, steps):
        return lib().obfv_rotate_keyswitch_count(self._c, steps)

    def behz_lift(self, poly_q):
        out = np.zeros((self.nbsk, self.N), dtype=np.uint64)
        lib().obfv_behz_lift(self._c, np.ascontiguousarray(poly_q), out)
        return out

    def behz_scale(self, in_q, in_bsk):
        out = np.zeros((self.L, self.N), dtype=np.uint64)
        lib().obfv_behz_scale(self._c, np.ascontiguousarray(in_q), np.ascontiguousarray(in_bsk), out)
        return out

    def switch_key(self, ct2, target, key):
        out = np.ascontiguousarray(ct2, dtype=np.uint64).copy()
        key = np.ascontiguousarray(key)
        lib().obfv_switch_key(self._c, out, np.ascontiguousarray(target), key.ctypes.data)
        return out
