"""Python host mirror of ABC's ciphertext boundary, over the C ABI (include/abc_b200.h).

`CudaCiphertextFactory` / `CudaCiphertext` carry the same method names, argument meaning and error
behaviour as the reference's SealCiphertextFactory / SealCiphertext
(/root/reference/include/ast_opt/runtime/SealCiphertextFactory.h:59-123,
 /root/reference/include/ast_opt/runtime/SealCiphertext.h:46-112), so parity tests read like the
reference's own (test/runtime/SealCiphertextFactoryTest.cpp).  The C++ drop-in that ABC's RuntimeVisitor
drives lives in abc_b200/cpp/ (CudaCiphertextFactory.h); this module is the binding used by tests and bench.py.

Everything here only sequences calls into libabc_b200.so — no arithmetic on ciphertext data in Python.
"""
import ctypes as C

import numpy as np

from . import _capi


def seal_parameters_from_bytes(data):
    """EncryptionParameters::load: -> dict(poly_degree, primes, plain_modulus) for CudaCiphertextFactory(...)."""
    lib = _capi.load()
    p = _capi.AbcParams()
    primes = (C.c_uint64 * 64)()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    if lib.abc_seal_params_parse(buf, len(data), C.byref(p), primes, 64) != 0:
        raise AbcError(lib.abc_last_error(None).decode())
    return dict(poly_degree=p.poly_degree, primes=[int(primes[i]) for i in range(p.n_primes)], plain_modulus=int(p.plain_modulus))

KEY_SECRET, KEY_PUBLIC, KEY_RELIN, KEY_GALOIS = 0, 1, 2, 3


class AbcError(RuntimeError):
    """std::runtime_error of the reference (every failure on this path is one:
    /root/reference/src/runtime/SealCiphertext.cpp:40,137,242)."""


class CudaCiphertextFactory:
    """Owns one device context: parameters, tables, keys, stream (SealCiphertextFactory.cpp:72-100)."""

    def __init__(self, numElementsPerCiphertextSlot=16384, primes=None, plain_modulus=0, device=0, batch=1,
                 seed=None, keygen=True, galois_steps=None):
        """seed=None: keys from the OS generator, like the reference's randomly seeded SEAL PRNG; an explicit seed
        makes the keys reproducible (tests, or the same keys on every GPU of a job)."""
        self._lib = _capi.load()
        p = _capi.AbcParams()
        p.poly_degree = numElementsPerCiphertextSlot
        self._primes_arr = None
        if primes:
            self._primes_arr = (C.c_uint64 * len(primes))(*primes)
            p.n_primes, p.primes = len(primes), self._primes_arr
        p.plain_modulus, p.device, p.batch, p.seed = plain_modulus, device, batch, seed or 0
        h = C.c_void_p()
        st = self._lib.abc_ctx_create(C.byref(p), C.byref(h))
        if st != 0:
            raise AbcError("abc_ctx_create failed (%d): %s" % (st, self._lib.abc_last_error(None).decode()))
        self._h = h
        self.N = self._lib.abc_poly_degree(h)
        self.k = self._lib.abc_n_primes(h)
        self.L = self._lib.abc_n_limbs(h)
        self.batch = self._lib.abc_batch(h)
        self.t = self._lib.abc_plain_modulus(h)
        q = np.zeros(self.k, dtype=np.uint64)
        self._ck(self._lib.abc_get_primes(h, q.ctypes.data))
        self.primes = [int(v) for v in q]
        if keygen:
            self.keygen(galois_steps)

    # -- plumbing
    def _ck(self, st):
        if st != 0:
            raise AbcError(self._lib.abc_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.abc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self._lib.abc_sync(self._h))

    def getCiphertextSlotSize(self):
        return self.N

    def aux_primes(self):
        out = np.zeros(self.L + 3, dtype=np.uint64)
        n = C.c_uint32()
        self._ck(self._lib.abc_get_aux_primes(self._h, out.ctypes.data, C.byref(n)))
        return int(out[0]), int(out[1]), [int(v) for v in out[2:n.value]]

    # -- keys
    def keygen(self, galois_steps=None):
        """Default: secret, public, relin and SEAL's default Galois key set.  galois_steps: only the keys for these
        rotation steps (large parameter sets, where one key is hundreds of MiB)."""
        if galois_steps is None:
            self._ck(self._lib.abc_keygen(self._h))
        else:
            elts = np.asarray([self.elt_from_step(s) for s in galois_steps], dtype=np.uint32)
            self._ck(self._lib.abc_keygen_select(self._h, elts.ctypes.data, elts.size))

    def elt_from_step(self, step):
        e = int(self._lib.abc_galois_elt_from_step(self._h, step))
        if e == 0:
            raise AbcError("step count too large")
        return e

    def key_shape(self, kind):
        return {KEY_SECRET: (self.k, self.N), KEY_PUBLIC: (2, self.k, self.N),
                KEY_RELIN: (self.L, 2, self.k, self.N), KEY_GALOIS: (self.L, 2, self.k, self.N)}[kind]

    def export_key(self, kind, galois_elt=0):
        out = np.zeros(self.key_shape(kind), dtype=np.uint64)
        self._ck(self._lib.abc_key_export(self._h, kind, galois_elt, out.ctypes.data, out.size))
        return out

    def import_key(self, kind, data, galois_elt=0):
        data = np.ascontiguousarray(data, dtype=np.uint64)
        self._ck(self._lib.abc_key_import(self._h, kind, galois_elt, data.ctypes.data, data.size))

    def has_galois_key(self, elt):
        return bool(self._lib.abc_has_galois_key(self._h, elt))

    # -- limb sharding across GPUs (one process per GPU; include/abc_b200.h "limb sharding")
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._ck(self._lib.abc_comm_unique_id(self._h, buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.abc_comm_init(self._h, rank, world, buf))

    def comm_stats(self):
        """(bytes received through all-gathers so far, NCCL collectives issued) on this rank."""
        b, n = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.abc_comm_stats(self._h, C.byref(b), C.byref(n)))
        return b.value, n.value

    def owned_limbs(self):
        lo, hi = C.c_uint32(), C.c_uint32()
        self._ck(self._lib.abc_owned_limbs(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def set_encrypt_nonce(self, nonce):
        self._ck(self._lib.abc_set_encrypt_nonce(self._h, nonce))

    def set_rng_key(self, key32):
        """The sampler's 32-byte ChaCha20 key from the caller (abc_set_rng_key): call on a factory built with
        keygen=False, then keygen()."""
        if len(key32) != 32:
            raise AbcError("rng key must be 32 bytes")
        buf = (C.c_ubyte * 32)(*key32)
        self._ck(self._lib.abc_set_rng_key(self._h, buf))

    # -- ciphertext creation (AbstractCiphertextFactory.h:19-38)
    def _slots(self, data):
        """Returns (int64 array, n per instance, broadcast flag)."""
        if np.isscalar(data):
            data = [data]
        arr = np.ascontiguousarray(data, dtype=np.int64)
        if arr.ndim == 1:
            if arr.size == 0:
                raise AbcError("Cannot encode an empty vector.")
            return arr, arr.size, 1
        if arr.ndim == 2 and arr.shape[0] == self.batch:
            return arr, arr.shape[1], 0
        raise AbcError("slot data must be 1-D (broadcast) or [batch][n]")

    def createCiphertext(self, data):
        arr, n, bc = self._slots(data)
        h = C.c_void_p()
        self._ck(self._lib.abc_encode_encrypt(self._h, arr.ctypes.data, n, bc, C.byref(h)))
        return CudaCiphertext(self, h)

    def createPlaintext(self, data):
        arr, n, bc = self._slots(data)
        h = C.c_void_p()
        self._ck(self._lib.abc_pt_encode(self._h, arr.ctypes.data, n, bc, C.byref(h)))
        return CudaPlaintext(self, h)

    def encryptPlaintext(self, pt):
        h = C.c_void_p()
        self._ck(self._lib.abc_encrypt_pt(self._h, pt._h, C.byref(h)))
        return CudaCiphertext(self, h)

    def allocCiphertext(self):
        h = C.c_void_p()
        self._ck(self._lib.abc_ct_alloc(self._h, C.byref(h)))
        return CudaCiphertext(self, h)

    def importCiphertext(self, words):
        """words: uint64 [batch][2][L][N] (per instance: seal::Ciphertext's coefficient layout)."""
        ct = self.allocCiphertext()
        words = np.ascontiguousarray(words, dtype=np.uint64)
        self._ck(self._lib.abc_ct_import(self._h, ct._h, words.ctypes.data, words.size))
        return ct

    def decryptCiphertext(self, ct):
        """Returns int64 [batch][N] ([N] when batch == 1), like decryptCiphertext fills its out-vector
        (SealCiphertextFactory.cpp:146-152)."""
        out = np.zeros((self.batch, self.N), dtype=np.int64)
        self._ck(self._lib.abc_decrypt_decode(self._h, ct._h, out.ctypes.data))
        return out[0] if self.batch == 1 else out

    def noiseBits(self, ct):
        """SealCiphertext::noiseBits (SealCiphertext.cpp:80-83): invariant noise budget, int32 [batch] (int when batch == 1)."""
        out = np.zeros(self.batch, dtype=np.int32)
        self._ck(self._lib.abc_noise_budget(self._h, ct._h, out.ctypes.data))
        return int(out[0]) if self.batch == 1 else out

    def getString(self, ct):
        vals = np.atleast_2d(self.decryptCiphertext(ct))[0]
        return "[" + ",".join(" %d" % v for v in vals) + " ]"

    # -- Microsoft SEAL 3.6 binary streams (include/abc_b200.h "SEAL 3.6 binary streams"; csrc/sealio.cu)
    def _save(self, call):
        n = C.c_size_t()
        call(None, 0, C.byref(n))                      # size query (fails with "buffer too small", fills n)
        buf = (C.c_uint8 * n.value)()
        self._ck(call(buf, n.value, C.byref(n)))
        return bytes(buf[:n.value])

    def sealParmsId(self, key_level=False):
        out = (C.c_uint64 * 4)()
        self._ck(self._lib.abc_seal_parms_id(self._h, int(key_level), out))
        return bytes(out)

    def saveSealParameters(self, compr=0):
        return self._save(lambda b, cap, n: self._lib.abc_seal_params_save(self._h, compr, b, cap, n))

    def saveSealCiphertext(self, ct, instance=0, compr=0):
        """seal::Ciphertext::save of one instance of the batch."""
        return self._save(lambda b, cap, n: self._lib.abc_seal_ct_save(self._h, ct._h, instance, compr, b, cap, n))

    def loadSealCiphertext(self, data, instance=0, into=None):
        """seal::Ciphertext::load into one instance of `into` (a new handle when omitted)."""
        ct = into if into is not None else self.allocCiphertext()
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self._ck(self._lib.abc_seal_ct_load(self._h, ct._h, instance, buf, len(data)))
        return ct

    def saveSealKey(self, kind, compr=0):
        return self._save(lambda b, cap, n: self._lib.abc_seal_key_save(self._h, kind, compr, b, cap, n))

    def loadSealKey(self, kind, data):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self._ck(self._lib.abc_seal_key_load(self._h, kind, buf, len(data)))

    def galois_elts(self):
        n = C.c_size_t()
        self._ck(self._lib.abc_galois_elts(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        self._ck(self._lib.abc_galois_elts(self._h, out.ctypes.data, out.size, C.byref(n)))
        return [int(e) for e in out]

    # -- probes / measurement
    def probe_ntt(self, mod_index, rows, inverse=False):
        rows = np.ascontiguousarray(rows, dtype=np.uint64).copy()
        r2 = rows.reshape(-1, self.N)
        self._ck(self._lib.abc_probe_ntt(self._h, int(inverse), mod_index, r2.ctypes.data, r2.shape[0]))
        return rows

    def probe_multiply(self, a, b):
        out = np.zeros((self.batch, 3, self.L, self.N), dtype=np.uint64)
        self._ck(self._lib.abc_probe_multiply(self._h, a._h, b._h, out.ctypes.data, out.size))
        return out

    def bench_ntt(self, mod_index, n_rows, iters=10, inverse=False):
        ms = C.c_float()
        self._ck(self._lib.abc_bench_ntt(self._h, int(inverse), mod_index, n_rows, iters, C.byref(ms)))
        return ms.value

    def timer_start(self):
        self._ck(self._lib.abc_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self._lib.abc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self, nbytes=256 << 20):
        self._ck(self._lib.abc_flush_l2(self._h, nbytes))

    def launch_count(self):
        return int(self._lib.abc_launch_count(self._h))

    def key_switch_count(self):
        return int(self._lib.abc_key_switch_count(self._h))

    def profile_enable(self, on=True):
        self._ck(self._lib.abc_profile_enable(self._h, int(on)))

    def profile(self):
        import json
        return json.loads(self._lib.abc_profile_json(self._h).decode())

    def measure_int_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self._lib.abc_measure_int_peak(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ntt_arith_class(self):
        return int(self._lib.abc_ntt_arith_class(self._h))

    def measure_butterfly_peak(self, arith_class=None):
        a = C.c_double()
        ar = self.ntt_arith_class() if arith_class is None else arith_class
        self._ck(self._lib.abc_measure_butterfly_peak(self._h, ar, C.byref(a)))
        return a.value


class CudaPlaintext:
    def __init__(self, factory, handle):
        self.factory, self._h = factory, handle

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self.factory, "_h", None):
            self.factory._lib.abc_pt_free(self._h)
        self._h = None


class CudaCiphertext:
    """Mirror of SealCiphertext (src/runtime/SealCiphertext.cpp): out-of-place ops return a new object and
    leave operands untouched; *Inplace ops overwrite self."""

    def __init__(self, factory, handle):
        self.factory, self._h = factory, handle

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self.factory, "_h", None):
            self.factory._lib.abc_ct_free(self._h)
        self._h = None

    def getFactory(self):
        return self.factory

    def noiseBits(self):
        return self.factory.noiseBits(self)

    def isTransparent(self):
        """Ciphertext::is_transparent per instance (SEAL throws std::logic_error on such results):
        bool [batch] (bool when batch == 1).  Synchronises."""
        f = self.factory
        out = np.zeros(f.batch, dtype=np.int32)
        f._ck(f._lib.abc_is_transparent(f._h, self._h, out.ctypes.data))
        return bool(out[0]) if f.batch == 1 else out.astype(bool)

    def _other(self, operand):
        if not isinstance(operand, CudaCiphertext) or operand.factory is not self.factory:
            raise AbcError("Cast of AbstractCiphertext to CudaCiphertext failed!")
        return operand

    def _new(self):
        return self.factory.allocCiphertext()

    def export(self):
        f = self.factory
        out = np.zeros((f.batch, 2, f.L, f.N), dtype=np.uint64)
        f._ck(f._lib.abc_ct_export(f._h, self._h, out.ctypes.data, out.size))
        return out

    def allgather(self):
        """Limb-sharded contexts: make every limb valid on every rank (NCCL all-gather)."""
        f = self.factory
        f._ck(f._lib.abc_ct_allgather(f._h, self._h))
        return self

    def clone(self):
        f = self.factory
        h = C.c_void_p()
        f._ck(f._lib.abc_ct_clone(f._h, self._h, C.byref(h)))
        return CudaCiphertext(f, h)

    # ctxt-ctxt
    def _cc(self, fn, operand, dst):
        f = self.factory
        f._ck(fn(f._h, dst._h, self._h, self._other(operand)._h))
        return dst

    def add(self, operand):
        return self._cc(self.factory._lib.abc_add, operand, self._new())

    def addInplace(self, operand):
        self._cc(self.factory._lib.abc_add, operand, self)

    def subtract(self, operand):
        return self._cc(self.factory._lib.abc_sub, operand, self._new())

    def subtractInplace(self, operand):
        self._cc(self.factory._lib.abc_sub, operand, self)

    def multiply(self, operand):
        return self._cc(self.factory._lib.abc_mul_relin, operand, self._new())

    def multiplyInplace(self, operand):
        self._cc(self.factory._lib.abc_mul_relin, operand, self)

    def negate(self):
        f, dst = self.factory, self._new()
        f._ck(f._lib.abc_negate(f._h, dst._h, self._h))
        return dst

    def negateInplace(self):
        f = self.factory
        f._ck(f._lib.abc_negate(f._h, self._h, self._h))

    def rotateRows(self, steps):
        f, dst = self.factory, self._new()
        f._ck(f._lib.abc_rotate_rows(f._h, dst._h, self._h, steps))
        return dst

    def rotateRowsAdd(self, steps, addend):
        """add(rotateRows(steps), addend) as one key switch (abc_rotate_rows_add).  The same fusion happens by itself
        when a rotateRows result is added (the library defers the last key switch of a rotation)."""
        f, dst = self.factory, self._new()
        f._ck(f._lib.abc_rotate_rows_add(f._h, dst._h, self._h, steps, self._other(addend)._h))
        return dst

    def isDeferred(self):
        return bool(self.factory._lib.abc_ct_deferred(self._h))

    def sharedCount(self):
        return int(self.factory._lib.abc_ct_shared(self._h))

    def rotateRowsInplace(self, steps):
        f = self.factory
        f._ck(f._lib.abc_rotate_rows(f._h, self._h, self._h, steps))

    # ctxt-plain; operand: list/array of ints (Cleartext<int>::getData()) or a CudaPlaintext
    def _cp(self, fn_slots, fn_pt, operand, dst):
        f = self.factory
        if isinstance(operand, CudaPlaintext):
            f._ck(fn_pt(f._h, dst._h, self._h, operand._h))
        else:
            arr, n, bc = f._slots(operand)
            f._ck(fn_slots(f._h, dst._h, self._h, arr.ctypes.data, n, bc))
        return dst

    @staticmethod
    def _all_minus_one(operand):
        if isinstance(operand, CudaPlaintext):
            return False
        a = np.asarray(operand)
        return a.size > 0 and bool((a == -1).all())

    def addPlain(self, operand):
        L = self.factory._lib
        return self._cp(L.abc_add_plain, L.abc_add_plain_pt, operand, self._new())

    def addPlainInplace(self, operand):
        L = self.factory._lib
        self._cp(L.abc_add_plain, L.abc_add_plain_pt, operand, self)

    def subtractPlain(self, operand):
        L = self.factory._lib
        return self._cp(L.abc_sub_plain, L.abc_sub_plain_pt, operand, self._new())

    def subtractPlainInplace(self, operand):
        L = self.factory._lib
        self._cp(L.abc_sub_plain, L.abc_sub_plain_pt, operand, self)

    def multiplyPlain(self, operand):
        # Cleartext<int>::allEqual(-1) -> negate fast path (SealCiphertext.cpp:156-157)
        if self._all_minus_one(operand):
            return self.negate()
        L = self.factory._lib
        return self._cp(L.abc_mul_plain, L.abc_mul_plain_pt, operand, self._new())

    def multiplyPlainInplace(self, operand):
        if self._all_minus_one(operand):
            return self.negateInplace()
        L = self.factory._lib
        self._cp(L.abc_mul_plain, L.abc_mul_plain_pt, operand, self)
