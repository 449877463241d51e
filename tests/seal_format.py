"""An independent Python restatement of Microsoft SEAL 3.6's binary stream format (seal::Serialization; ciphertext.cpp,
plaintext.cpp, publickey.h, secretkey.h, kswitchkeys.cpp, encryptionparams.cpp, modulus.cpp, dynarray.h), used only by
the tests to write streams for the product to load and to parse the streams the product saves.  It shares no code with
abc_b200/csrc/sealio.cu: struct / zlib / hashlib.blake2b from the standard library.  SEAL itself is not installable
here, so neither side has seen a byte written by real SEAL ("parity unpinned")."""
import hashlib
import struct
import zlib

import numpy as np

MAGIC, HEADER = 0xA15E, 16
NONE, ZLIB, ZSTD = 0, 1, 2


def parms_id(N, primes, t):
    """EncryptionParameters::compute_parms_id: BLAKE2b-256 over [scheme = bfv (1), N, q_0 .. q_{k-1}, t] as u64."""
    words = [1, N] + list(primes) + [t]
    return hashlib.blake2b(struct.pack("<%dQ" % len(words), *words), digest_size=32).digest()


def record(body, compr=NONE):
    if compr == ZLIB:
        body = zlib.compress(body)
    elif compr != NONE:
        raise ValueError("compr_mode")
    return struct.pack("<HBBBBHQ", MAGIC, HEADER, 3, 6, compr, 0, HEADER + len(body)) + body


def open_record(buf, at=0, nested=False):
    """-> (body bytes, offset after the record)"""
    magic, hs, major, minor, compr, _, size = struct.unpack_from("<HBBBBHQ", buf, at)
    assert magic == MAGIC and hs == HEADER and major == 3 and minor == 6, "header"
    assert HEADER <= size <= len(buf) - at, "size field"
    body = bytes(buf[at + HEADER:at + size])
    if compr == ZLIB:
        assert not nested
        body = zlib.decompress(body)
    else:
        assert compr == NONE, "compr_mode %d" % compr
    return body, at + size


def dynarray(words):
    words = np.ascontiguousarray(words, dtype="<u8").ravel()
    return record(struct.pack("<Q", words.size) + words.tobytes())


def ciphertext_body(pid, ntt, data, N, cms, size=2):
    return pid + struct.pack("<BQQQd", int(ntt), size, N, cms, 1.0) + dynarray(data)


def ciphertext(N, data_primes, t_and_all, data, compr=NONE):
    """seal::Ciphertext at the data level; data uint64 [2][L][N]; t_and_all = (t, key-level primes)."""
    t, allp = t_and_all
    return record(ciphertext_body(parms_id(N, data_primes, t), False, data, N, len(data_primes)), compr)


def parse_ciphertext_body(body, at=0):
    pid = body[at:at + 32]
    ntt, size, N, cms, scale = struct.unpack_from("<BQQQd", body, at + 32)
    arr, end = open_record(body, at + 32 + 33, nested=True)
    (count,) = struct.unpack_from("<Q", arr, 0)
    words = np.frombuffer(arr, dtype="<u8", offset=8, count=count)
    assert len(arr) == 8 + 8 * count and count == size * N * cms
    return dict(parms_id=pid, ntt=bool(ntt), size=size, N=N, cms=cms, scale=scale,
                data=words.reshape(size, cms, N).copy()), end


def parse_ciphertext(buf):
    body, _ = open_record(buf)
    return parse_ciphertext_body(body)[0]


def secret_key(N, primes, t, data, compr=NONE):
    """SecretKey = record{ Plaintext record{parms_id, coeff_count, scale, DynArray} }; data uint64 [k][N] (NTT form)."""
    k = len(primes)
    pt = parms_id(N, primes, t) + struct.pack("<Qd", k * N, 1.0) + dynarray(data)
    return record(record(pt), compr)


def parse_secret_key(buf):
    body, _ = open_record(buf)
    pt, _ = open_record(body, 0, nested=True)
    cc, scale = struct.unpack_from("<Qd", pt, 32)
    arr, _ = open_record(pt, 48, nested=True)
    (count,) = struct.unpack_from("<Q", arr, 0)
    assert count == cc
    return dict(parms_id=pt[:32], data=np.frombuffer(arr, dtype="<u8", offset=8, count=count).copy())


def public_key(N, primes, t, data, compr=NONE):
    """PublicKey = record{ Ciphertext record (NTT form, key level) }; data uint64 [2][k][N]."""
    return record(record(ciphertext_body(parms_id(N, primes, t), True, data, N, len(primes))), compr)


def parse_public_key(buf):
    body, _ = open_record(buf)
    ct, _ = open_record(body, 0, nested=True)
    return parse_ciphertext_body(ct)[0]


def kswitch_keys(N, primes, t, keys, dim1, compr=NONE):
    """KSwitchKeys::save_members: parms_id, dim1, then per index dim2 + that many PublicKey records.
    keys: {index: uint64 [L][2][k][N]}; RelinKeys: {0: ...}, dim1 = 1; GaloisKeys: index = (elt - 1) // 2, dim1 = N."""
    pid = parms_id(N, primes, t)
    out = [pid, struct.pack("<Q", dim1)]
    for idx in range(dim1):
        if idx not in keys:
            out.append(struct.pack("<Q", 0))
            continue
        kd = keys[idx]
        out.append(struct.pack("<Q", kd.shape[0]))
        for J in range(kd.shape[0]):
            out.append(record(record(ciphertext_body(pid, True, kd[J], N, len(primes)))))
    return record(b"".join(out), compr)


def parse_kswitch_keys(buf):
    body, _ = open_record(buf)
    pid = body[:32]
    (dim1,) = struct.unpack_from("<Q", body, 32)
    at, keys = 40, {}
    for idx in range(dim1):
        (dim2,) = struct.unpack_from("<Q", body, at)
        at += 8
        rows = []
        for _ in range(dim2):
            pk, at = open_record(body, at, nested=True)
            ct, _ = open_record(pk, 0, nested=True)
            rows.append(parse_ciphertext_body(ct)[0])
        if rows:
            assert all(r["parms_id"] == pid and r["ntt"] for r in rows)
            keys[idx] = np.stack([r["data"] for r in rows])
    assert at == len(body)
    return dict(parms_id=pid, dim1=dim1, keys=keys)


def encryption_parameters(N, primes, t, compr=NONE):
    body = struct.pack("<BQQ", 1, N, len(primes))
    for q in list(primes) + [t]:
        body += record(struct.pack("<Q", q))
    return record(body, compr)


def parse_encryption_parameters(buf):
    body, _ = open_record(buf)
    scheme, N, k = struct.unpack_from("<BQQ", body, 0)
    at, vals = 17, []
    for _ in range(k + 1):
        m, at = open_record(body, at, nested=True)
        vals.append(struct.unpack("<Q", m)[0])
    assert at == len(body)
    return dict(scheme=scheme, N=N, primes=vals[:-1], t=vals[-1])
