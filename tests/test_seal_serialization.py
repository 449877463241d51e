"""SEAL 3.6 binary streams (include/abc_b200.h "SEAL 3.6 binary streams", abc_b200/csrc/sealio.cu) against an independent
Python restatement of the format (tests/seal_format.py).  CPU part: the restatement against itself and fixed facts of
the format; GPU part: streams written in Python from ORACLE ciphertexts / keys are loaded by the product and used
(decrypt, multiply, rotate: bit-exact vs the oracle), and streams saved by the product parse back to the same words."""
import ctypes as C
import hashlib
import struct

import numpy as np
import pytest

import seal_format as sf

SEED = 4673838


def test_format_restatement_round_trip():
    N, primes, t = 4096, [0xffffee001, 0xffffc4001, 0x1ffffe0001], 1032193
    rng = np.random.default_rng(0)
    data = rng.integers(0, primes[0], size=(2, 2, N), dtype=np.uint64)
    for compr in (sf.NONE, sf.ZLIB):
        buf = sf.ciphertext(N, primes[:2], (t, primes), data, compr)
        magic, hs, major, minor, cm, _, size = struct.unpack_from("<HBBBBHQ", buf, 0)
        assert (magic, hs, major, minor, cm, size) == (0xA15E, 16, 3, 6, compr, len(buf))
        got = sf.parse_ciphertext(buf)
        assert np.array_equal(got["data"], data) and not got["ntt"] and got["size"] == 2 and got["cms"] == 2
        assert got["parms_id"] == sf.parms_id(N, primes[:2], t)
    # an uncompressed ciphertext: 16-byte header, 32 + 1 + 3*8 + 8 bytes of metadata, a 16 + 8 byte DynArray prologue
    assert len(sf.ciphertext(N, primes[:2], (t, primes), data)) == 16 + 65 + 24 + data.size * 8
    p = sf.parse_encryption_parameters(sf.encryption_parameters(N, primes, t, sf.ZLIB))
    assert p == dict(scheme=1, N=N, primes=primes, t=t)
    keys = {0: rng.integers(0, 1 << 30, size=(2, 2, 3, N), dtype=np.uint64)}
    back = sf.parse_kswitch_keys(sf.kswitch_keys(N, primes, t, keys, 1))
    assert back["dim1"] == 1 and np.array_equal(back["keys"][0], keys[0])


def test_c_abi_rejects_garbage_without_a_device():
    """abc_seal_params_parse needs no context (and no GPU): header validation and the zlib path."""
    from abc_b200 import AbcError, seal_parameters_from_bytes
    N, primes, t = 8192, [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001], 1032193
    for compr in (sf.NONE, sf.ZLIB):
        got = seal_parameters_from_bytes(sf.encryption_parameters(N, primes, t, compr))
        assert got == dict(poly_degree=N, primes=primes, plain_modulus=t)
    good = sf.encryption_parameters(N, primes, t)
    for bad in (b"", good[:10], b"\x00" + good[1:], good[:3] + b"\x04" + good[4:], good[:-9],
                good[:8] + struct.pack("<Q", len(good) + 1) + good[16:]):
        with pytest.raises(AbcError):
            seal_parameters_from_bytes(bad)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def pair(oracle4096):
    from abc_b200 import CudaCiphertextFactory
    f = CudaCiphertextFactory(4096, seed=SEED)
    yield f, oracle4096
    f.close()


def _zstd():
    try:
        return C.CDLL("libzstd.so.1")
    except OSError:
        return None


@pytest.mark.gpu
def test_parms_id_is_blake2b_of_the_parameters(pair):
    f, o = pair
    assert f.sealParmsId(key_level=True) == sf.parms_id(f.N, f.primes, f.t)
    assert f.sealParmsId(key_level=False) == sf.parms_id(f.N, f.primes[:-1], f.t)
    words = [1, f.N] + f.primes + [f.t]
    assert f.sealParmsId(True) == hashlib.blake2b(struct.pack("<%dQ" % len(words), *words), digest_size=32).digest()
    assert sf.parse_encryption_parameters(f.saveSealParameters(sf.ZLIB)) == dict(scheme=1, N=f.N, primes=f.primes, t=f.t)


@pytest.mark.gpu
def test_load_ciphertexts_written_in_seal_format(pair):
    f, o = pair
    rng = np.random.default_rng(1)
    d1, d2 = rng.integers(0, 1025, f.N), rng.integers(0, 1025, f.N)
    w1, w2 = o.encrypt_slots(d1, 3), o.encrypt_slots(d2, 4)
    a = f.loadSealCiphertext(sf.ciphertext(f.N, f.primes[:-1], (f.t, f.primes), w1))
    b = f.loadSealCiphertext(sf.ciphertext(f.N, f.primes[:-1], (f.t, f.primes), w2, sf.ZLIB))
    assert np.array_equal(a.export()[0], w1) and np.array_equal(b.export()[0], w2)
    assert np.array_equal(f.decryptCiphertext(a), o.decrypt_slots(w1))
    assert np.array_equal(a.multiply(b).export()[0], o.mul_relin(w1, w2))
    assert np.array_equal(a.rotateRows(7).export()[0], o.rotate_rows(w1, 7))
    for compr in (sf.NONE, sf.ZLIB):                                    # and back out
        got = sf.parse_ciphertext(f.saveSealCiphertext(a, 0, compr))
        assert np.array_equal(got["data"], w1) and got["parms_id"] == sf.parms_id(f.N, f.primes[:-1], f.t)
        assert not got["ntt"] and got["scale"] == 1.0
    z = _zstd()
    if z is not None:                                                   # zstd: SEAL's default compr_mode when built with it
        raw = f.saveSealCiphertext(a, 0, sf.ZSTD)
        assert raw[5] == 2 and struct.unpack_from("<Q", raw, 8)[0] == len(raw)
        assert np.array_equal(f.loadSealCiphertext(raw).export()[0], w1)
        body = sf.ciphertext_body(sf.parms_id(f.N, f.primes[:-1], f.t), False, w2, f.N, f.L)
        z.ZSTD_compressBound.restype = C.c_size_t; z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        cap = z.ZSTD_compressBound(len(body))
        dst = C.create_string_buffer(cap)
        n = z.ZSTD_compress(dst, cap, body, len(body), 3)
        stream = struct.pack("<HBBBBHQ", 0xA15E, 16, 3, 6, 2, 0, 16 + n) + dst.raw[:n]
        assert np.array_equal(f.loadSealCiphertext(stream).export()[0], w2)


@pytest.mark.gpu
def test_load_rejects_mismatched_or_damaged_streams(pair):
    from abc_b200 import AbcError
    f, o = pair
    w = o.encrypt_slots([1, 2, 3], 5)
    good = sf.ciphertext(f.N, f.primes[:-1], (f.t, f.primes), w)
    other_parms = sf.ciphertext(f.N, f.primes[:-1], (f.t + 2, f.primes), w)           # different plain modulus
    too_big = w.copy(); too_big[1, 0, 5] = f.primes[0]                                # coefficient == q_0
    ntt_flag = sf.record(sf.ciphertext_body(sf.parms_id(f.N, f.primes[:-1], f.t), True, w, f.N, f.L))
    seeded = sf.record(sf.parms_id(f.N, f.primes[:-1], f.t) + struct.pack("<BQQQd", 0, 2, f.N, f.L, 1.0) + sf.dynarray(w[0]))
    for bad in (good[:100], good[:-8], other_parms, sf.ciphertext(f.N, f.primes[:-1], (f.t, f.primes), too_big),
                ntt_flag, seeded, sf.public_key(f.N, f.primes, f.t, o.public_key())):
        with pytest.raises(AbcError):
            f.loadSealCiphertext(bad)
    with pytest.raises(AbcError):
        f.loadSealCiphertext(good, instance=1)                                        # batch is 1


@pytest.mark.gpu
def test_keys_through_seal_streams(oracle4096):
    """A second context without keys receives every key as SEAL streams written from the ORACLE's keys, then
    computes bit-exactly; the product's own saved streams parse back to the same words."""
    from abc_b200 import KEY_GALOIS, KEY_PUBLIC, KEY_RELIN, KEY_SECRET, CudaCiphertextFactory
    o = oracle4096
    f = CudaCiphertextFactory(4096, seed=99, keygen=False)
    N, P, t = f.N, f.primes, f.t
    f.loadSealKey(KEY_SECRET, sf.secret_key(N, P, t, o.secret_key()))
    f.loadSealKey(KEY_PUBLIC, sf.public_key(N, P, t, o.public_key(), sf.ZLIB))
    f.loadSealKey(KEY_RELIN, sf.kswitch_keys(N, P, t, {0: o.relin_key()}, 1, sf.ZLIB))
    elts = o.galois_elts()
    f.loadSealKey(KEY_GALOIS, sf.kswitch_keys(N, P, t, {(e - 1) // 2: o.galois_key(e) for e in elts}, N))
    assert sorted(f.galois_elts()) == sorted(set(elts))
    rng = np.random.default_rng(2)
    d = rng.integers(0, 1025, N)
    w = o.encrypt_slots(d, 9)
    a = f.loadSealCiphertext(sf.ciphertext(N, P[:-1], (t, P), w))
    assert np.array_equal(f.decryptCiphertext(a), o.decrypt_slots(w))
    assert np.array_equal(a.multiply(a).export()[0], o.mul_relin(w, w))
    for k in (1, -3, 20):
        assert np.array_equal(a.rotateRows(k).export()[0], o.rotate_rows(w, k))
    # product -> streams -> independent parser
    sk = sf.parse_secret_key(f.saveSealKey(KEY_SECRET))
    assert np.array_equal(sk["data"].reshape(f.k, N), o.secret_key()) and sk["parms_id"] == sf.parms_id(N, P, t)
    pk = sf.parse_public_key(f.saveSealKey(KEY_PUBLIC, sf.ZLIB))
    assert np.array_equal(pk["data"], o.public_key()) and pk["ntt"] and pk["cms"] == f.k
    rk = sf.parse_kswitch_keys(f.saveSealKey(KEY_RELIN))
    assert rk["dim1"] == 1 and np.array_equal(rk["keys"][0], o.relin_key())
    gk = sf.parse_kswitch_keys(f.saveSealKey(KEY_GALOIS, sf.ZLIB))
    assert gk["dim1"] == N and sorted(gk["keys"]) == sorted({(e - 1) // 2 for e in elts})
    for e in elts:
        assert np.array_equal(gk["keys"][(e - 1) // 2], o.galois_key(e))
    f.close()
