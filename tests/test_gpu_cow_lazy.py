"""Copy-on-write clones and deferred rotations (include/abc_b200.h: abc_ct_clone, abc_rotate_rows,
abc_rotate_rows_add) must be invisible in the results: every sequence below is compared bit for bit with the
oracle's eager evaluation.  RuntimeVisitor clones on every variable read (src/runtime/RuntimeVisitor.cpp:436) and
adds rotated variables (`acc = acc +++ r`), which is what these two mechanisms exist for."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 4673838


@pytest.fixture(scope="module", params=["4096", "8192"])
def pair(request, oracle4096, oracle8192):
    from abc_b200 import CudaCiphertextFactory
    n = int(request.param)
    f = CudaCiphertextFactory(n, seed=SEED)
    yield f, (oracle4096 if n == 4096 else oracle8192)
    f.close()


def fresh(f, o, rng, nonce):
    d = rng.integers(0, 1025, size=f.N, dtype=np.int64)
    f.set_encrypt_nonce(nonce)
    return f.createCiphertext(d), o.encrypt_slots(d, nonce)


def eq(ct, words):
    return np.array_equal(ct.export()[0], words)


def test_clone_is_shared_until_written(pair):
    f, o = pair
    rng = np.random.default_rng(5)
    a, aw = fresh(f, o, rng, 31)
    b, bw = fresh(f, o, rng, 32)
    c = a.clone()
    assert a.sharedCount() == 2 and c.sharedCount() == 2
    c.addInplace(b)                       # writes c only
    assert a.sharedCount() == 1 and c.sharedCount() == 1
    assert eq(a, aw) and eq(c, o.add(aw, bw))
    d = a.clone()
    a.multiplyInplace(b)                  # writes a; the clone keeps the old value
    assert eq(d, aw) and eq(a, o.mul_relin(aw, bw))
    e = d.clone()
    e.subtractPlainInplace([3, 4, 5])
    assert eq(d, aw) and eq(e, o.sub_plain(aw, o.encode(o.expand([3, 4, 5]))))
    g = d.clone()
    g.rotateRowsInplace(3)
    assert eq(d, aw) and eq(g, o.rotate_rows(aw, 3))
    del c, d, e, g


@pytest.mark.parametrize("steps", [1, -24, 63, 5])   # direct keys and NAF paths (63 = 64 - 1, 5 = 4 + 1)
def test_deferred_rotation_then_add(pair, steps):
    f, o = pair
    rng = np.random.default_rng(6)
    a, aw = fresh(f, o, rng, 41)
    b, bw = fresh(f, o, rng, 42)
    rw = o.rotate_rows(aw, steps)
    r = a.rotateRows(steps)
    assert r.isDeferred()
    s1 = b.add(r)                         # fused; r now stands for the difference s1 - b, still nothing stored
    assert eq(s1, o.add(bw, rw))
    assert r.isDeferred()
    n0 = f.launch_count()
    assert eq(r, rw) and not r.isDeferred()
    assert f.launch_count() - n0 == 1     # one subtraction, not a second key switch
    assert eq(b.add(r), o.add(bw, rw))    # and r is an ordinary ciphertext from here on
    r2 = a.rotateRows(steps)
    s2 = r2.add(b)                        # deferred operand on the left
    s3 = r2.add(a)                        # the difference node used in another add
    assert eq(s2, o.add(rw, bw)) and eq(s3, o.add(rw, aw)) and eq(r2, rw)
    acc = b.clone()
    r4 = a.rotateRows(steps)
    acc.addInplace(r4)                    # RuntimeVisitor's `acc = acc +++ r`: the addend is also the destination
    acc.addInplace(a)                     # the sum buffer is still referenced by r4: copy-on-write keeps r4 intact
    assert eq(acc, o.add(o.add(bw, rw), aw)) and eq(r4, rw) and eq(b, bw)
    r3 = a.rotateRows(steps)
    r3.addInplace(b)                      # the only holder is overwritten: single-output path
    assert eq(r3, o.add(rw, bw))
    assert eq(a, aw) and eq(b, bw)


def test_deferred_rotation_other_consumers(pair):
    f, o = pair
    rng = np.random.default_rng(7)
    a, aw = fresh(f, o, rng, 51)
    b, bw = fresh(f, o, rng, 52)
    rw = o.rotate_rows(aw, 2)
    assert eq(a.rotateRows(2), rw)                                        # export resolves
    assert np.array_equal(f.decryptCiphertext(a.rotateRows(2)), o.decrypt_slots(rw))
    assert eq(a.rotateRows(2).multiply(b), o.mul_relin(rw, bw))
    assert eq(b.subtract(a.rotateRows(2)), o.sub(bw, rw))
    assert eq(a.rotateRows(2).rotateRows(-7), o.rotate_rows(rw, -7))      # rotation of a deferred rotation
    assert eq(a.rotateRows(2).addPlain([9, 8, 7]), o.add_plain(rw, o.encode(o.expand([9, 8, 7]))))
    r = a.rotateRows(2)
    assert eq(r.add(r), o.add(rw, rw))                                    # both operands the same deferred buffer
    r1, r2 = a.rotateRows(2), b.rotateRows(-1)
    assert eq(r1.add(r2), o.add(rw, o.rotate_rows(bw, -1)))               # two deferred operands
    c = a.rotateRows(2).clone()                                           # clone of a deferred handle
    assert eq(c, rw)
    z = a.rotateRows(0)
    assert not z.isDeferred() and eq(z, aw)
    a2 = a.clone()
    a2.rotateRowsInplace(2)
    a2.addInplace(a)                                                      # s = rot(s) + s, in place
    assert eq(a2, o.add(rw, aw)) and eq(a, aw)


def test_explicit_rotate_add_and_ladder(pair):
    f, o = pair
    rng = np.random.default_rng(8)
    a, aw = fresh(f, o, rng, 61)
    b, bw = fresh(f, o, rng, 62)
    assert eq(a.rotateRowsAdd(-24, b), o.add(o.rotate_rows(aw, -24), bw))
    assert eq(a.rotateRowsAdd(0, b), o.add(aw, bw))
    s, sw = a, aw                                                         # the bench's rotate-and-sum ladder
    for k in (8, 4, 2, 1):
        s, sw = s.add(s.rotateRows(k)), o.add(sw, o.rotate_rows(sw, k))
    assert eq(s, sw)
    assert np.array_equal(f.decryptCiphertext(s), o.decrypt_slots(sw))


def test_multiply_by_itself_takes_squaring_path(pair):
    """multiply(x, x) lifts and transforms one operand only (lib.cu behz_multiply); same coefficients as the
    general path, which the oracle always takes."""
    f, o = pair
    rng = np.random.default_rng(9)
    a, aw = fresh(f, o, rng, 71)
    want = o.mul_relin(aw, aw)
    assert eq(a.multiply(a), want)
    assert eq(a.multiply(a.clone()), want)          # clones share the buffer: also the squaring path
    assert np.array_equal(f.probe_multiply(a, a)[0], o.multiply(aw, aw))   # the size-3 product before relinearisation
    b = a.clone()
    b.multiplyInplace(b)
    assert eq(b, want) and eq(a, aw)


def test_noise_budget_matches_oracle(pair):
    """abc_noise_budget = Decryptor::invariant_noise_budget (SealCiphertext::noiseBits, SealCiphertext.cpp:80-83)."""
    f, o = pair
    rng = np.random.default_rng(10)
    a, aw = fresh(f, o, rng, 81)
    assert a.noiseBits() == o.noise_budget(aw) > 0
    r = a.rotateRows(5)
    assert r.noiseBits() == o.noise_budget(o.rotate_rows(aw, 5))
    m, mw = a, aw
    seen = []
    for _ in range(5):
        m, mw = m.multiply(m), o.mul_relin(mw, mw)
        seen.append(m.noiseBits())
        assert seen[-1] == o.noise_budget(mw)
    assert seen[0] < a.noiseBits() and seen[-1] == 0


def test_rotation_prefix_cache_saves_key_switches_and_never_goes_stale(pair, monkeypatch):
    """rotate(x, 63) = step -1 then step 64 (SEAL's NAF order, low digit first), rotate(x, -65) = step -1 then step -64,
    and rotate(x, -1) is that first step itself: the library keeps the intermediate buffer per (source buffer, Galois
    element), so the eight rotations of a 3x3 stencil cost 8 key switches instead of 12 — with the same coefficients — and a
    source that is written afterwards (copy-on-write: the cache holds a reference) never yields a stale result."""
    from abc_b200 import CudaCiphertextFactory
    f, o = pair
    rng = np.random.default_rng(9)
    a, aw = fresh(f, o, rng, 71)
    b, bw = fresh(f, o, rng, 72)
    steps = (-65, -64, -63, -1, 1, 63, 64, 65)

    def stencil(fac, img):
        acc = None
        for k in steps:
            r = img.rotateRows(k)
            if acc is None:
                acc = r
            else:
                acc.addInplace(r)
        return acc

    want = None
    for k in steps:
        r = o.rotate_rows(aw, k)
        want = r if want is None else o.add(want, r)
    n0 = f.key_switch_count()
    got = stencil(f, a)
    assert eq(got, want)
    with_cache = f.key_switch_count() - n0
    monkeypatch.setenv("ABC_ROT_CACHE", "0")              # the same without the cache
    h = CudaCiphertextFactory(f.N, seed=SEED)
    try:
        a2 = h.importCiphertext(a.export())
        n1 = h.key_switch_count()
        got2 = stencil(h, a2)
        without = h.key_switch_count() - n1
        assert np.array_equal(got2.export(), got.export())
    finally:
        h.close()
    assert (with_cache, without) == (8, 12)
    # staleness: the cached prefix belongs to the OLD buffer of `a`
    a.addInplace(b)                                                   # writes a (a new buffer: the cache pinned the old one)
    aw2 = o.add(aw, bw)
    assert eq(a.rotateRows(63), o.rotate_rows(aw2, 63))
    assert eq(a.rotateRows(-1), o.rotate_rows(aw2, -1))
    c = a.clone()
    c.rotateRowsInplace(-65)                                          # shares a's cached prefix (same buffer), then diverges
    assert eq(c, o.rotate_rows(aw2, -65)) and eq(a, aw2)
    a.negateInplace()
    assert eq(a.rotateRows(63), o.rotate_rows(o.negate(aw2), 63))
