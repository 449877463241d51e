"""Random straight-line programs over the ciphertext interface, coefficient-exact against the oracle.

The op-by-op parity tests pin every operation on fresh operands; RuntimeVisitor, however, produces arbitrary interleavings of
clones (src/runtime/RuntimeVisitor.cpp:436), in-place and out-of-place ops, and rotations whose key switch the library defers
to the consumer (include/abc_b200.h: abc_rotate_rows).  Here a seeded generator draws such sequences — operands may alias
(x op x, dst = operand), handles are dropped at random, deferred handles are consumed by every kind of op — and every live
variable is compared word for word with the oracle's eager evaluation at random points and at the end.  Results need not
be decryptable (the oracle does the same modular arithmetic whatever the noise), so depth is not limited."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 4673838
STEPS = (1, -1, 2, 4, -24, 63, 5, -7, 1024, -2047)      # direct keys and NAF chains


def _run_program(f, o, seed, n_ops, batch):
    rng = np.random.default_rng(seed)
    live = []                                             # (handle, [oracle words per instance])

    def fresh(nonce):
        d = rng.integers(-500, 501, size=(batch, 16), dtype=np.int64)
        f.set_encrypt_nonce(nonce)
        ct = f.createCiphertext(d if batch > 1 else d[0])
        return ct, [o.encrypt_slots(d[i], nonce * batch + i) for i in range(batch)]

    def check(idx, what):
        ct, words = live[idx]
        got = ct.export()
        for i in range(batch):
            assert np.array_equal(got[i], words[i]), "seed %d: variable %d instance %d differs after %s" % (seed, idx, i, what)

    def each(fn, *operands):
        return [fn(*[w[i] for w in operands]) for i in range(batch)]

    for n in range(3):
        live.append(fresh(100 + n))
    trace = []
    for step in range(n_ops):
        kind = rng.choice(["add", "sub", "mul", "rot", "neg", "addp", "subp", "mulp", "clone", "drop", "peek", "rotadd"],
                          p=[.16, .10, .07, .20, .04, .05, .04, .06, .10, .06, .06, .06])
        a = int(rng.integers(len(live)))
        b = int(rng.integers(len(live)))                   # may equal a: x op x
        inplace = bool(rng.integers(2))
        ca, wa = live[a]
        cb, wb = live[b]
        trace.append((kind, a, b, inplace))
        if kind in ("add", "sub", "mul"):
            ofn = {"add": o.add, "sub": o.sub, "mul": o.mul_relin}[kind]
            res_w = each(ofn, wa, wb)
            if inplace:
                getattr(ca, {"add": "addInplace", "sub": "subtractInplace", "mul": "multiplyInplace"}[kind])(cb)
                live[a] = (ca, res_w)
            else:
                live.append((getattr(ca, {"add": "add", "sub": "subtract", "mul": "multiply"}[kind])(cb), res_w))
        elif kind == "rot":
            s = int(rng.choice(STEPS))
            res_w = each(lambda w: o.rotate_rows(w, s), wa)
            if inplace:
                ca.rotateRowsInplace(s); live[a] = (ca, res_w)
            else:
                live.append((ca.rotateRows(s), res_w))     # stays deferred until something consumes it
        elif kind == "rotadd":
            s = int(rng.choice(STEPS))
            live.append((ca.rotateRowsAdd(s, cb), each(lambda x, y: o.add(o.rotate_rows(x, s), y), wa, wb)))
        elif kind == "neg":
            res_w = each(o.negate, wa)
            if inplace:
                ca.negateInplace(); live[a] = (ca, res_w)
            else:
                live.append((ca.negate(), res_w))
        elif kind in ("addp", "subp", "mulp"):
            vals = [int(v) for v in rng.integers(-9, 10, size=int(rng.integers(1, 6)))]
            if kind == "mulp" and all(v == 0 for v in vals):
                vals[0] = 3
            plain = o.encode(o.expand(vals))
            ofn = {"addp": o.add_plain, "subp": o.sub_plain, "mulp": o.multiply_plain}[kind]
            res_w = each(lambda w: ofn(w, plain), wa)
            name = {"addp": "addPlain", "subp": "subtractPlain", "mulp": "multiplyPlain"}[kind]
            if inplace:
                getattr(ca, name + "Inplace")(vals); live[a] = (ca, res_w)
            else:
                live.append((getattr(ca, name)(vals), res_w))
        elif kind == "clone":
            live.append((ca.clone(), [w.copy() for w in wa]))
        elif kind == "drop" and len(live) > 3:
            del ca, cb
            live.pop(a)
        elif kind == "peek":
            check(a, "op %d (%s)" % (step, trace[-6:]))
        if len(live) > 10:                                 # bound the pool: drop the oldest
            live.pop(0)
    for idx in range(len(live)):
        check(idx, "the whole program")
    # and one decryption through the device path against the oracle's
    ct, words = live[-1]
    dec = np.atleast_2d(f.decryptCiphertext(ct))
    for i in range(batch):
        assert np.array_equal(dec[i], o.decrypt_slots(words[i]))


@pytest.mark.parametrize("seed", range(12))
def test_random_programs_n4096(seed, oracle4096):
    from abc_b200 import CudaCiphertextFactory
    batch = 1 + seed % 3
    f = CudaCiphertextFactory(4096, seed=SEED, batch=batch)
    try:
        _run_program(f, oracle4096, 1000 + seed, 100, batch)
    finally:
        f.close()


@pytest.mark.parametrize("seed,env", [(0, {}), (1, {"ABC_EAGER_ROTATE": "1"}), (2, {"ABC_KS_ONE_LAUNCH": "0"})])
def test_random_programs_n8192(seed, env, oracle8192, monkeypatch):
    from abc_b200 import CudaCiphertextFactory
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    f = CudaCiphertextFactory(8192, seed=SEED, batch=2)
    try:
        _run_program(f, oracle8192, 2000 + seed, 40, 2)
    finally:
        f.close()


def test_random_program_n16384_default_size():
    """ABC's default factory size: wide exact-double class, half-limb key-switch rows, BEHZ14."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    o = Oracle(16384, seed=SEED)
    f = CudaCiphertextFactory(16384, seed=SEED)
    try:
        _run_program(f, o, 3000, 24, 1)
    finally:
        f.close()
