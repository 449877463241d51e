"""GPU parity: every CUDA path against the CPU oracle, bit-exact on ciphertext coefficients, through the
C ABI (abc_b200._capi -> libabc_b200.so).  Shapes follow the reference's
test/runtime/SealCiphertextFactoryTest.cpp (N=4096) plus the BASELINE N=8192 / 16384 parameter sets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 4673838


@pytest.fixture(scope="module")
def pair4096(oracle4096):
    from abc_b200 import CudaCiphertextFactory
    f = CudaCiphertextFactory(4096, seed=SEED)
    yield f, oracle4096
    f.close()


@pytest.fixture(scope="module")
def pair8192(oracle8192):
    from abc_b200 import CudaCiphertextFactory
    f = CudaCiphertextFactory(8192, seed=SEED)
    yield f, oracle8192
    f.close()


def rand_slots(rng, n, lo=0, hi=1025):
    return rng.integers(lo, hi, size=n, dtype=np.int64)


def test_parameters_match_oracle(pair4096, pair8192):
    for f, o in (pair4096, pair8192):
        assert f.primes == o.primes and f.t == o.t and f.k == o.k and f.L == o.L
        assert f.aux_primes() == o.aux_primes()


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_ntt_forward_inverse_bit_exact(which, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(1)
    nmods = f.k + (f.L + 1) + 1
    for mi in range(nmods):
        if mi < f.k:
            q, oidx = f.primes[mi], mi
        elif mi < f.k + f.L + 1:
            msk, _, B = o.aux_primes()
            q, oidx = (B + [msk])[mi - f.k], 1000 + mi - f.k
        else:
            q, oidx = f.t, -1
        rows = rng.integers(0, q, size=(3, f.N), dtype=np.uint64)
        got = f.probe_ntt(mi, rows)
        for r in range(3):
            assert np.array_equal(got[r], o.ntt_fwd(oidx, rows[r])), "forward NTT mod index %d" % mi
        back = f.probe_ntt(mi, got, inverse=True)
        assert np.array_equal(back, rows), "inverse NTT mod index %d" % mi


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_keygen_bit_exact(which, pair4096, pair8192):
    from abc_b200 import KEY_GALOIS, KEY_PUBLIC, KEY_RELIN, KEY_SECRET
    f, o = pair4096 if which == "4096" else pair8192
    assert np.array_equal(f.export_key(KEY_SECRET), o.secret_key())
    assert np.array_equal(f.export_key(KEY_PUBLIC), o.public_key())
    assert np.array_equal(f.export_key(KEY_RELIN), o.relin_key())
    elts = o.galois_elts()
    assert len(set(elts)) == 2 * int(np.log2(f.N)) - 2
    for e in elts:
        assert f.has_galois_key(e)
        assert np.array_equal(f.export_key(KEY_GALOIS, e), o.galois_key(e)), "galois key %d" % e


def test_caller_supplied_rng_key_matches_oracle():
    """abc_set_rng_key: the sampler's 32-byte ChaCha20 key from the caller (the oracle's key stream is pinned to RFC 8439 in
    tests/test_oracle.py); keys and an encryption under it are word for word the oracle's, and differ under another key."""
    from abc_b200 import CudaCiphertextFactory, KEY_GALOIS, KEY_PUBLIC, KEY_RELIN, KEY_SECRET
    from oracle.bfv_oracle import Oracle
    key = bytes((37 * i + 11) % 256 for i in range(32))
    o = Oracle(4096, rng_key=key, galois_steps=[1, -2])
    f = CudaCiphertextFactory(4096, keygen=False, batch=2)
    g = CudaCiphertextFactory(4096, keygen=False)
    try:
        f.set_rng_key(key)
        f.keygen([1, -2])
        assert np.array_equal(f.export_key(KEY_SECRET), o.secret_key())
        assert np.array_equal(f.export_key(KEY_PUBLIC), o.public_key())
        assert np.array_equal(f.export_key(KEY_RELIN), o.relin_key())
        for step in (1, -2):
            e = f.elt_from_step(step)
            assert np.array_equal(f.export_key(KEY_GALOIS, e), o.galois_key(e))
        data = np.arange(2 * 9).reshape(2, 9) - 4
        f.set_encrypt_nonce(21)
        ct = f.createCiphertext(data).export()
        for i in range(2):
            assert np.array_equal(ct[i], o.encrypt_slots(data[i], 21 * 2 + i))
        g.set_rng_key(bytes(32))
        g.keygen([1, -2])
        assert not np.array_equal(g.export_key(KEY_SECRET), o.secret_key())
    finally:
        f.close(); g.close()


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_encrypt_decrypt_bit_exact(which, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(2)
    for nonce, n in ((11, 6), (12, f.N), (13, 1)):
        data = rand_slots(rng, n, -1000, 1000)
        f.set_encrypt_nonce(nonce)
        ct = f.createCiphertext(data)
        want = o.encrypt_slots(data, nonce)
        assert np.array_equal(ct.export()[0], want)
        dec = f.decryptCiphertext(ct)
        assert np.array_equal(dec, o.decrypt_slots(want))
        assert np.array_equal(dec, o.expand(data))


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_add_sub_negate_bit_exact(which, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(3)
    a_w, b_w = o.encrypt_slots(rand_slots(rng, f.N), 21), o.encrypt_slots(rand_slots(rng, f.N), 22)
    a, b = f.importCiphertext(a_w[None]), f.importCiphertext(b_w[None])
    assert np.array_equal(a.add(b).export()[0], o.add(a_w, b_w))
    assert np.array_equal(a.subtract(b).export()[0], o.sub(a_w, b_w))
    assert np.array_equal(a.negate().export()[0], o.negate(a_w))
    # operands unchanged (SealCiphertextFactoryTest.cpp:157-159)
    assert np.array_equal(a.export()[0], a_w) and np.array_equal(b.export()[0], b_w)
    a.addInplace(b)
    assert np.array_equal(a.export()[0], o.add(a_w, b_w))


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_multiply_and_relinearize_bit_exact(which, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(4)
    da, db = rand_slots(rng, f.N), rand_slots(rng, f.N)
    a_w, b_w = o.encrypt_slots(da, 31), o.encrypt_slots(db, 32)
    a, b = f.importCiphertext(a_w[None]), f.importCiphertext(b_w[None])
    m3 = o.multiply(a_w, b_w)
    assert np.array_equal(f.probe_multiply(a, b)[0], m3), "BEHZ size-3 product"
    want = o.relinearize(m3)
    got = a.multiply(b)
    assert np.array_equal(got.export()[0], want), "mul+relin"
    assert np.array_equal(f.decryptCiphertext(got), (da * db) % f.t - f.t * (((da * db) % f.t) > f.t // 2))
    a.multiplyInplace(b)
    assert np.array_equal(a.export()[0], want)


@pytest.mark.parametrize("which", ["4096", "8192"])
@pytest.mark.parametrize("steps", [1, 4, -24, 63, -1, 6])
def test_rotate_rows_bit_exact(which, steps, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(5)
    d = rand_slots(rng, f.N)
    a_w = o.encrypt_slots(d, 41)
    a = f.importCiphertext(a_w[None])
    got = a.rotateRows(steps)
    assert np.array_equal(got.export()[0], o.rotate_rows(a_w, steps))
    half = f.N // 2
    want = np.concatenate([np.roll(d[:half], -steps), np.roll(d[half:], -steps)])
    assert np.array_equal(f.decryptCiphertext(got), want)
    assert np.array_equal(a.export()[0], a_w)
    a.rotateRowsInplace(steps)
    assert np.array_equal(a.export()[0], o.rotate_rows(a_w, steps))


def test_rotate_step_out_of_range_raises(pair4096):
    from abc_b200 import AbcError
    f, _ = pair4096
    ct = f.createCiphertext([1, 2, 3])
    with pytest.raises(AbcError):
        ct.rotateRows(f.N // 2)
    assert np.array_equal(f.decryptCiphertext(ct.rotateRows(0)), f.decryptCiphertext(ct))


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_plain_ops_bit_exact(which, pair4096, pair8192):
    f, o = pair4096 if which == "4096" else pair8192
    rng = np.random.default_rng(6)
    a_w = o.encrypt_slots(rand_slots(rng, f.N), 51)
    a = f.importCiphertext(a_w[None])
    for data in ([0, 1, 2, 1, 10, 21], list(rand_slots(rng, f.N, -500, 500)), [19], [-7]):
        pl = o.encode(o.expand(data))
        assert np.array_equal(a.addPlain(data).export()[0], o.add_plain(a_w, pl))
        assert np.array_equal(a.subtractPlain(data).export()[0], o.sub_plain(a_w, pl))
        assert np.array_equal(a.multiplyPlain(data).export()[0], o.multiply_plain(a_w, pl))
    # all-(-1) fast path = negate (SealCiphertext.cpp:156-157)
    assert np.array_equal(a.multiplyPlain([-1, -1, -1]).export()[0], o.negate(a_w))
    b = a.clone()
    b.multiplyPlainInplace([3])
    assert np.array_equal(b.export()[0], o.multiply_plain(a_w, o.encode(o.expand([3]))))
    assert np.array_equal(a.export()[0], a_w)


def test_reference_kats_through_cuda_factory(pair4096):
    """Slot-level vectors of test/runtime/SealCiphertextFactoryTest.cpp:44-336 through the CUDA factory."""
    import json, os
    f, _ = pair4096
    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "abc_kats.json")))

    def check(ct, expected):
        got = f.decryptCiphertext(ct)
        assert got.shape == (f.N,)
        assert list(got[:len(expected)]) == expected
        assert (got[len(expected):] == expected[-1]).all()

    v = kats["factory"]
    check(f.createCiphertext(v["create"]), v["create"])
    a, b = f.createCiphertext(v["a"]), f.createCiphertext(v["b"])
    check(a.add(b), v["add"]); check(a.subtract(b), v["sub"]); check(a.multiply(b), v["mul"])
    check(a, v["a"]); check(b, v["b"])
    check(a.addPlain(v["b"]), v["add"]); check(a.subtractPlain(v["b"]), v["sub"]); check(a.multiplyPlain(v["b"]), v["mul"])
    for op, key in (("addInplace", "add"), ("subtractInplace", "sub"), ("multiplyInplace", "mul")):
        c = a.clone(); getattr(c, op)(b); check(c, v[key])
    for op, key in (("addPlainInplace", "add"), ("subtractPlainInplace", "sub"), ("multiplyPlainInplace", "mul")):
        c = a.clone(); getattr(c, op)(v["b"]); check(c, v[key])
    # rotation vectors (:51-140)
    d = v["rotate_data"]
    half = f.N // 2
    full = np.array(d + [d[-1]] * (f.N - len(d)), dtype=np.int64)
    ct = f.createCiphertext(d)
    for steps in v["rotate_steps"]:
        want = np.concatenate([np.roll(full[:half], -steps), np.roll(full[half:], -steps)])
        assert np.array_equal(f.decryptCiphertext(ct.rotateRows(steps)), want)
    check(ct, d)


def test_too_many_elements_raises(pair4096):
    from abc_b200 import AbcError
    f, _ = pair4096
    with pytest.raises(AbcError):
        f.createCiphertext(list(range(f.N + 1)))


def test_batch_instances_independent(oracle4096):
    """batch=3: each instance equals the single-instance oracle result with nonce*B+b."""
    from abc_b200 import CudaCiphertextFactory
    o = oracle4096
    f = CudaCiphertextFactory(4096, seed=SEED, batch=3)
    try:
        rng = np.random.default_rng(7)
        da, db = rand_slots(rng, (3, 100)), rand_slots(rng, (3, 100))
        f.set_encrypt_nonce(5)
        a = f.createCiphertext(da)
        b = f.createCiphertext(db)
        a_w = [o.encrypt_slots(da[i], 5 * 3 + i) for i in range(3)]
        b_w = [o.encrypt_slots(db[i], 6 * 3 + i) for i in range(3)]
        ea, prod, rot = a.export(), a.multiply(b).export(), a.rotateRows(-24).export()
        for i in range(3):
            assert np.array_equal(ea[i], a_w[i])
            assert np.array_equal(prod[i], o.mul_relin(a_w[i], b_w[i]))
            assert np.array_equal(rot[i], o.rotate_rows(a_w[i], -24))
        dec = f.decryptCiphertext(a.add(b))
        for i in range(3):
            assert np.array_equal(dec[i], o.expand(da[i] + db[i]))
    finally:
        f.close()


def test_n16384_default_parameters():
    """ABC's default factory size (SealCiphertextFactory.h:13: 16384 slots, k=9)."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    o = Oracle(16384, seed=SEED)
    f = CudaCiphertextFactory(16384, seed=SEED)
    try:
        assert f.primes == o.primes and f.t == o.t == 786433
        rng = np.random.default_rng(8)
        da, db = rand_slots(rng, f.N), rand_slots(rng, f.N)
        f.set_encrypt_nonce(3)
        a = f.createCiphertext(da)
        a_w, b_w = o.encrypt_slots(da, 3), o.encrypt_slots(db, 4)
        assert np.array_equal(a.export()[0], a_w)
        b = f.importCiphertext(b_w[None])
        assert np.array_equal(a.multiply(b).export()[0], o.mul_relin(a_w, b_w))
        assert np.array_equal(a.rotateRows(63).export()[0], o.rotate_rows(a_w, 63))
        assert np.array_equal(f.decryptCiphertext(a.multiply(b)), (da * db + f.t // 2) % f.t - f.t // 2)
    finally:
        f.close()


@pytest.fixture(scope="module")
def pair32768():
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    o = Oracle(32768, seed=SEED)
    f = CudaCiphertextFactory(32768, seed=SEED)
    yield f, o
    f.close()


def test_n32768_two_pass_ntt_bit_exact(pair32768):
    """N = 32768 (k = 16, 55-bit primes): head passes in registers + 8192-coefficient tail blocks in shared memory."""
    f, o = pair32768
    assert f.primes == o.primes and f.t == o.t == 786433 and f.k == 16
    rng = np.random.default_rng(11)
    msk, _, B = o.aux_primes()
    for mi, q, oidx in ((0, f.primes[0], 0), (15, f.primes[15], 15), (16, B[0], 1000), (16 + 15, msk, 1015),
                        (16 + 16, f.t, -1)):
        rows = rng.integers(0, q, size=(2, f.N), dtype=np.uint64)
        got = f.probe_ntt(mi, rows)
        assert np.array_equal(got[0], o.ntt_fwd(oidx, rows[0])) and np.array_equal(got[1], o.ntt_fwd(oidx, rows[1]))
        assert np.array_equal(f.probe_ntt(mi, got, inverse=True), rows)


def test_n32768_ops_bit_exact(pair32768):
    from abc_b200 import KEY_PUBLIC, KEY_RELIN, KEY_SECRET
    f, o = pair32768
    assert np.array_equal(f.export_key(KEY_SECRET), o.secret_key())
    assert np.array_equal(f.export_key(KEY_PUBLIC), o.public_key())
    assert np.array_equal(f.export_key(KEY_RELIN), o.relin_key())
    rng = np.random.default_rng(12)
    da, db = rand_slots(rng, f.N), rand_slots(rng, f.N)
    f.set_encrypt_nonce(7)
    a = f.createCiphertext(da)
    a_w, b_w = o.encrypt_slots(da, 7), o.encrypt_slots(db, 8)
    assert np.array_equal(a.export()[0], a_w)
    b = f.importCiphertext(b_w[None])
    assert np.array_equal(f.decryptCiphertext(a), da)
    assert np.array_equal(a.multiply(b).export()[0], o.mul_relin(a_w, b_w))
    for steps in (1, -24):
        assert np.array_equal(a.rotateRows(steps).export()[0], o.rotate_rows(a_w, steps))
    pl = o.encode(o.expand([5, -3, 7]))
    assert np.array_equal(a.multiplyPlain([5, -3, 7]).export()[0], o.multiply_plain(a_w, pl))
    assert np.array_equal(a.addPlain([5, -3, 7]).export()[0], o.add_plain(a_w, pl))
    assert np.array_equal(f.decryptCiphertext(a.multiply(b)), (da * db + f.t // 2) % f.t - f.t // 2)


@pytest.mark.parametrize("which", ["4096", "8192"])
def test_behz_product_extreme_and_random_residues(which, pair4096, pair8192):
    """The BEHZ product as a function of raw residues (not of valid encryptions): all q_i - 1, all zero, alternating
    0 / q_i - 1, a single non-zero coefficient, and uniformly random residues — the inputs that stretch the ranges of the
    base conversions.  The CUDA path (on the exact-double class: its own sub-2^45 auxiliary base, csrc/behz_f64.cuh)
    must equal the oracle (SEAL's 61-bit base) bit for bit, also when both operands are the same handle (squaring)."""
    f, o = pair4096 if which == "4096" else pair8192
    L, N = f.L, f.N
    q = np.array(f.primes[:L], dtype=np.uint64)
    rng = np.random.default_rng(99)

    def fill(kind):
        ct = np.zeros((2, L, N), dtype=np.uint64)
        for i in range(L):
            if kind == "max":
                ct[:, i, :] = q[i] - np.uint64(1)
            elif kind == "alt":
                ct[:, i, ::2] = q[i] - np.uint64(1)
            elif kind == "one":
                ct[0, i, 0] = q[i] - np.uint64(1); ct[1, i, N - 1] = np.uint64(1)
            elif kind == "rand":
                ct[:, i, :] = rng.integers(0, int(q[i]), size=(2, N), dtype=np.uint64)
        return ct

    cases = [("max", "max"), ("max", "rand"), ("zero", "rand"), ("alt", "one"), ("rand", "rand"), ("one", "max")]
    for ka, kb in cases:
        a_w, b_w = fill(ka), fill(kb)
        a, b = f.importCiphertext(a_w[None]), f.importCiphertext(b_w[None])
        assert np.array_equal(f.probe_multiply(a, b)[0], o.multiply(a_w, b_w)), "product %s x %s" % (ka, kb)
        assert np.array_equal(f.probe_multiply(a, a)[0], o.multiply(a_w, a_w)), "square %s" % ka
