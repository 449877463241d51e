"""CPU tests: the oracle against the reference's known-answer vectors and algebraic identities
(SURVEY.md App. A.9).  No GPU."""
import json
import os

import numpy as np
import pytest

from oracle.bfv_oracle import Oracle, get_primes

SEED = 4673838

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "abc_kats.json")))


def check(o, ct, expected):
    got = o.decrypt_slots(ct)
    assert list(got[:len(expected)]) == expected
    assert (got[len(expected):] == expected[-1]).all()  # pad-with-last (SealCiphertextFactoryTest.cpp:38-40)


def test_default_parameters():
    """SEAL CoeffModulus::BFVDefault / PlainModulus::Batching(N,20)."""
    assert get_primes(4096, 20, 1) == [1032193] and get_primes(8192, 20, 1) == [1032193]
    assert get_primes(16384, 20, 1) == [786433] and get_primes(32768, 20, 1) == [786433]
    o = Oracle(8192)
    assert o.primes == [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001]
    msk, gamma, B = o.aux_primes()
    assert (msk, gamma) == (0x1ffffffffffa4001, 0x1ffffffffff74001)
    assert B == [0x1ffffffffff0c001, 0x1fffffffffec4001, 0x1fffffffffe10001, 0x1fffffffffe00001]
    for N, bits in ((4096, 109), (8192, 218), (16384, 438)):
        q = Oracle(N).primes
        assert sum(p.bit_length() for p in q) == bits
        assert all((p - 1) % (2 * N) == 0 for p in q)


def test_minimal_root_is_minimal(oracle4096):
    o = oracle4096
    q, N = o.primes[0], o.N
    psi = o.psi(0)
    assert pow(psi, N, q) == q - 1
    roots = [pow(psi, e, q) for e in range(1, 2 * N, 2)]
    assert psi == min(roots)


def test_ntt_is_negacyclic_evaluation(oracle4096):
    """out[bitrev(i)] = a(psi^(2i+1)); INTT(NTT(a)) = a; dyadic product = negacyclic convolution."""
    o = oracle4096
    q, N = o.primes[1], o.N
    rng = np.random.default_rng(0)
    a = rng.integers(0, q, size=N, dtype=np.uint64)
    A = o.ntt_fwd(1, a)
    assert np.array_equal(o.ntt_inv(1, A), a)
    psi, logn = o.psi(1), N.bit_length() - 1
    for i in (0, 1, 5, N - 1):
        x = pow(psi, 2 * i + 1, q)
        val, p = 0, 1
        for c in a:
            val = (val + int(c) * p) % q
            p = p * x % q
        assert int(A[int(format(i, "0%db" % logn)[::-1], 2)]) == val
    # sparse negacyclic product check: a * X^3 * 2
    b = np.zeros(N, dtype=np.uint64); b[3] = 2
    prod = o.ntt_inv(1, (A.astype(object) * o.ntt_fwd(1, b).astype(object) % q).astype(np.uint64))
    want = np.roll(a.astype(object) * 2 % q, 3)
    want[:3] = (q - want[:3]) % q
    assert np.array_equal(prod.astype(object), want)


def test_encoder_roundtrip_and_slot_semantics(oracle4096):
    o = oracle4096
    rng = np.random.default_rng(1)
    v = rng.integers(-(o.t // 2), o.t // 2 + 1, size=o.N, dtype=np.int64)
    assert np.array_equal(o.decode(o.encode(v)), v)
    # a constant vector encodes to a constant polynomial (broadcast scalars)
    pl = o.encode(np.full(o.N, 19, dtype=np.int64))
    assert pl[0] == 19 and not pl[1:].any()


def test_factory_kats(oracle4096):
    o, v = oracle4096, KATS["factory"]
    check(o, o.encrypt_slots(v["create"], 1), v["create"])
    a, b = o.encrypt_slots(v["a"], 2), o.encrypt_slots(v["b"], 3)
    check(o, o.add(a, b), v["add"])
    check(o, o.sub(a, b), v["sub"])
    check(o, o.mul_relin(a, b), v["mul"])
    pb = o.encode(o.expand(v["b"]))
    check(o, o.add_plain(a, pb), v["add"])
    check(o, o.sub_plain(a, pb), v["sub"])
    check(o, o.multiply_plain(a, pb), v["mul"])
    # size-3 product decrypts to the same slots before relinearisation
    check(o, o.multiply(a, b), v["mul"])


def test_rotation_kats(oracle4096):
    """SealCiphertextFactoryTest.cpp:51-140 incl. wrap-around at N/2."""
    o, v = oracle4096, KATS["factory"]
    d, half = v["rotate_data"], o.N // 2
    ct = o.encrypt_slots(d, 4)
    for steps in v["rotate_steps"]:
        dv = o.decrypt_slots(o.rotate_rows(ct, steps))
        for i in range(o.N):
            if steps > 0:
                if i < min(len(d) - steps, half - steps):
                    assert dv[i] == d[i + steps]
                elif half - steps <= i < half:
                    assert dv[i] == d[i - (half - steps)]
                else:
                    assert dv[i] == d[-1]
            else:
                if i < -steps or i >= -steps + len(d):
                    assert dv[i] == d[-1]
                else:
                    assert dv[i] == d[i + steps]
    with pytest.raises(ValueError):
        o.rotate_rows(ct, half)


def test_runtime_visitor_kats(oracle4096):
    o = oracle4096
    for n, kat in enumerate(KATS["runtime_visitor"]):
        ct = o.encrypt_slots(kat["input"], 100 + n)
        for op in kat["ops"]:
            if op[0] == "rotate":
                ct = o.rotate_rows(ct, op[1])
            elif op[0] == "mul_ct2":
                ct = o.mul_relin(ct, o.encrypt_slots(kat["input2"], 200 + n))
            elif op[0] == "mul_plain":
                ct = o.multiply_plain(ct, o.encode(o.expand(op[1])))
            elif op[0] == "accumulate_self":
                acc = o.encrypt_slots([0], 300 + n)
                for _ in range(op[1]):
                    acc = o.add(acc, ct)
                ct = acc
        got = o.decrypt_slots(ct)
        assert list(got[:len(kat["expect"])]) == kat["expect"], kat["name"]


def test_default_galois_key_set(oracle4096):
    o = oracle4096
    elts = o.galois_elts()
    assert len(set(elts)) == 2 * 12 - 2
    assert [o.rotate_keyswitch_count(s) for s in (1, 2, 4, -4, -1, 1024, 2047)] == [1] * 7
    assert [o.rotate_keyswitch_count(s) for s in (-24, 6, 63, 3)] == [2, 2, 2, 2]


def test_galois_ntt_table_matches_coefficient_automorphism(oracle4096):
    """Galois key for elt decrypts sigma_elt(c1): key switch with it turns (sigma(c0), sigma(c1)) into a valid ct."""
    o = oracle4096
    rng = np.random.default_rng(3)
    d = rng.integers(0, 1025, size=o.N, dtype=np.int64)
    ct = o.encrypt_slots(d, 9)
    for step in (1, -1, 512):
        got = o.decrypt_slots(o.apply_galois(ct, o.elt_from_step(step)))
        half = o.N // 2
        assert np.array_equal(got, np.concatenate([np.roll(d[:half], -step), np.roll(d[half:], -step)]))
    # column swap (elt 2N-1)
    got = o.decrypt_slots(o.apply_galois(ct, 2 * o.N - 1))
    assert np.array_equal(got, np.concatenate([d[o.N // 2:], d[:o.N // 2]]))


def test_homomorphic_identities_random(oracle8192):
    """decrypt(op(encrypt)) on random slots incl. negatives and values near t/2 (App. A.9 item 4), N=8192."""
    o = oracle8192
    rng = np.random.default_rng(4)
    t = o.t
    x = rng.integers(-(t // 2), t // 2 + 1, size=o.N, dtype=np.int64)
    y = rng.integers(-(t // 2), t // 2 + 1, size=o.N, dtype=np.int64)
    cx, cy = o.encrypt(o.encode(x), 1), o.encrypt(o.encode(y), 2)

    def centre(v):
        v = v % t
        return np.where(v > t // 2, v - t, v)

    assert np.array_equal(o.decrypt_slots(cx), x)
    assert np.array_equal(o.decrypt_slots(o.add(cx, cy)), centre(x + y))
    assert np.array_equal(o.decrypt_slots(o.sub(cx, cy)), centre(x - y))
    assert np.array_equal(o.decrypt_slots(o.negate(cx)), centre(-x))
    assert np.array_equal(o.decrypt_slots(o.mul_relin(cx, cy)), centre(x * y))
    assert np.array_equal(o.decrypt_slots(o.multiply_plain(cx, o.encode(y))), centre(x * y))
    assert np.array_equal(o.decrypt_slots(o.add_plain(cx, o.encode(y))), centre(x + y))
    # noise budget sanity: three sequential mul+relin survive at N=8192 (App. A.9 item 6)
    small = rng.integers(0, 4, size=o.N, dtype=np.int64)
    c = o.encrypt(o.encode(small), 3)
    want = small.copy()
    for _ in range(3):
        c = o.mul_relin(c, c)
        want = centre(want * want)
    assert np.array_equal(o.decrypt_slots(c), want)


def test_expand_rejects_oversize(oracle4096):
    with pytest.raises(RuntimeError):
        oracle4096.expand(list(range(4097)))


def test_sampler_distributions(oracle4096):
    """sk ternary, noise small: the only SEAL-side properties of the sampler that matter."""
    o = oracle4096
    q0 = o.primes[0]
    sk = o.ntt_inv(0, o.secret_key()[0])
    vals = np.where(sk > q0 // 2, sk.astype(np.int64) - q0, sk.astype(np.int64))
    assert set(np.unique(vals)) <= {-1, 0, 1}
    assert 0.25 < np.mean(vals == 0) < 0.42
    # pk: c0 + c1*s = -e, small
    pk = o.public_key()
    s = o.secret_key()[0].astype(object)
    e = o.ntt_inv(0, ((pk[0, 0].astype(object) + pk[1, 0].astype(object) * s) % q0).astype(np.uint64))
    ev = np.where(e > q0 // 2, e.astype(np.int64) - q0, e.astype(np.int64))
    assert np.abs(ev).max() <= 21 and 2.5 < ev.std() < 4.0


def test_sampler_key_stream_is_chacha20_rfc8439():
    """The sampler's randomness is the ChaCha20 key stream: RFC 8439 section 2.3.2 (key 00..1f, nonce 00:00:00:09 00:00:00:4a
    00:00:00:00, block counter 1).  In the sampler's terms (DESIGN.md "Sampler") the nonce words are (domain | b << 4, a lo,
    a hi) and word idx of a stream sits in block idx >> 3, so that block is words 8..15 of stream (domain 0, a = 0x4a000000,
    b = 0x09000000 >> 4)."""
    import ctypes as C
    from oracle.bfv_oracle import lib
    key = (C.c_ubyte * 32)(*range(32))
    block = bytes.fromhex(
        "10f1e7e4d13b5915500fdd1fa32071c4" "c7d1f4c733c068030422aa9ac3d46c4e"
        "d2826446079faa0914c2d705d98b02a2" "b5129cd1de164eb9cbd083e8a2503c4e")
    want = np.frombuffer(block, dtype="<u8")
    got = [lib().obfv_rng(0, key, 0, 0x4a000000, 0x09000000 >> 4, 8 + j) for j in range(8)]
    assert got == [int(v) for v in want]
    # a 64-bit seed expands to the key (seed lo, seed hi, "abc-", "b200", 0, 0, 0, 0)
    seed = 0x1122334455667788
    key2 = (C.c_ubyte * 32)(*(seed.to_bytes(8, "little") + b"abc-b200" + bytes(16)))
    assert lib().obfv_rng(seed, None, 4, 77, 1, 5) == lib().obfv_rng(0, key2, 4, 77, 1, 5)
    # distinct streams, distinct words
    assert len({lib().obfv_rng(seed, None, d, a, b, 0) for d in (1, 2, 3, 4) for a in (0, 1) for b in (0, 1, 2)}) == 24


def test_explicit_rng_key_replaces_the_seed():
    key = bytes(range(100, 132))
    a = Oracle(4096, rng_key=key, galois_steps=[1])
    b = Oracle(4096, rng_key=key, seed=999, galois_steps=[1])       # the seed no longer matters
    c = Oracle(4096, seed=0, galois_steps=[1])
    assert np.array_equal(a.secret_key(), b.secret_key()) and np.array_equal(a.public_key(), b.public_key())
    assert not np.array_equal(a.secret_key(), c.secret_key())


def test_noise_budget_tracks_decryptability(oracle4096):
    """invariant_noise_budget (SealCiphertext::noiseBits): positive while decryption is exact, shrinking with every
    multiplication, 0 once the noise has overrun q/t (N=4096: 72-bit q, 20-bit t, two multiplications fit)."""
    o = oracle4096
    rng = np.random.default_rng(3)
    d = rng.integers(0, 4, size=o.N, dtype=np.int64)
    ct = o.encrypt_slots(d, 5)
    fresh = o.noise_budget(ct)
    assert 25 <= fresh <= 50
    assert o.noise_budget(o.add(ct, ct)) in (fresh, fresh - 1, fresh - 2)
    want, budgets = d.copy(), [fresh]
    for _ in range(4):
        ct = o.mul_relin(ct, ct)
        want = (want * want) % o.t
        budgets.append(o.noise_budget(ct))
        ok = np.array_equal(o.decrypt_slots(ct) % o.t, want)
        assert ok or budgets[-1] == 0, "a positive budget guarantees decryption"
    assert budgets[1] < budgets[0] and budgets[-1] == 0
    assert all(a >= b for a, b in zip(budgets, budgets[1:]))


def test_behz_product_does_not_depend_on_the_auxiliary_base():
    """bfv_multiply lifts q -> Bsk = B U {m_sk}, multiplies, and comes back (fast_floor + Shenoy-Kumaresan).  Every
    conversion is exact or has an error term that depends on the q residues only, so WHICH primes make up B and m_sk must
    not show in the product as long as their product passes SEAL's size rule (more than 32 + bits(t) + bits(Q) bits).
    The CUDA path relies on this to run the product over 44-bit auxiliary primes on the FP64 pipe
    (abc_b200/csrc/behz_f64.cuh); here the oracle itself is run with SEAL's base and with such bases (test hook
    obfv_create_aux), on fresh encryptions, on a chain of three products, and on extreme raw residues.  A base below the
    size rule must (and does) give a different product, which shows the comparison is sensitive."""
    from oracle.bfv_oracle import Oracle
    N = 4096
    ref = Oracle(N, seed=SEED)
    L = ref.L
    q = np.array(ref.primes[:L], dtype=np.uint64)
    rng = np.random.default_rng(3)
    a = ref.encrypt_slots(rng.integers(0, 1025, N), 1)
    b = ref.encrypt_slots(rng.integers(0, 1025, N), 2)
    mx = np.zeros((2, L, N), dtype=np.uint64)
    rnd = np.zeros((2, L, N), dtype=np.uint64)
    for i in range(L):
        mx[:, i, :] = q[i] - np.uint64(1)
        rnd[:, i, :] = rng.integers(0, int(q[i]), size=(2, N), dtype=np.uint64)

    def products(o):
        m1 = o.multiply(a, b)
        r1 = o.relinearize(m1)
        m2 = o.multiply(r1, r1)
        m3 = o.multiply(o.relinearize(m2), a)
        return [m1, m2, m3, o.multiply(mx, mx), o.multiply(mx, rnd), o.multiply(rnd, rnd)]

    want = products(ref)
    # N = 4096: Q has 72 bits, t 20: the rule asks for more than 124 bits; 4 x 44 = 176 (what the CUDA path takes), 3 x 45, 5 x 40
    for aux in ((44, 3), (45, 2), (40, 4)):
        o = Oracle(N, seed=SEED, aux=aux)
        assert o.nbsk == aux[1] + 1
        for x, y in zip(want, products(o)):
            assert np.array_equal(x, y), "auxiliary base %r changes the product" % (aux,)
    small = Oracle(N, seed=SEED, aux=(44, 1))      # 88 bits: below the rule
    assert not all(np.array_equal(x, y) for x, y in zip(want, products(small)))


def test_oracle_regression_digests():
    """tests/golden/oracle_vectors.json: coefficient-level digests of the oracle's own outputs (keys and encryption under
    the ChaCha20 sampler, every op).  Pins the oracle against accidental change; says nothing about SEAL."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(here, "golden", "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "golden", "oracle_vectors.json")))
    for n in (4096, 8192):
        assert mod.vectors(n) == want[str(n)], "oracle output changed at N = %d" % n
