"""GPU: edge cases the reference tests or SEAL's API define for this path, and non-default parameter sets
(BASELINE.json configs[2]: 3-16 RNS limbs), all bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 4673838


def make_pair(N, primes=None, t=0):
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    o = Oracle(N, primes=primes, t=t, seed=SEED)
    f = CudaCiphertextFactory(N, primes=primes, plain_modulus=t, seed=SEED)
    return f, o


def check_ops(f, o, seed=0):
    rng = np.random.default_rng(seed)
    da, db = rng.integers(0, 1025, f.N), rng.integers(0, 1025, f.N)
    a_w, b_w = o.encrypt_slots(da, 1), o.encrypt_slots(db, 2)
    f.set_encrypt_nonce(1)
    a = f.createCiphertext(da)
    assert np.array_equal(a.export()[0], a_w), "encrypt"
    b = f.importCiphertext(b_w[None])
    assert np.array_equal(a.add(b).export()[0], o.add(a_w, b_w)), "add"
    assert np.array_equal(a.multiply(b).export()[0], o.mul_relin(a_w, b_w)), "mul+relin"
    for steps in (1, -24):
        assert np.array_equal(a.rotateRows(steps).export()[0], o.rotate_rows(a_w, steps)), "rotate %d" % steps
    pl = o.encode(o.expand([7, -2, 5]))
    assert np.array_equal(a.multiplyPlain([7, -2, 5]).export()[0], o.multiply_plain(a_w, pl)), "mul plain"
    assert np.array_equal(a.subtractPlain([7, -2, 5]).export()[0], o.sub_plain(a_w, pl)), "sub plain"
    assert np.array_equal(f.decryptCiphertext(a.multiply(b)), o.decrypt_slots(o.mul_relin(a_w, b_w))), "decrypt"


@pytest.mark.parametrize("N,bits,k", [(4096, 40, 3), (4096, 50, 4), (8192, 50, 7), (8192, 58, 3), (16384, 45, 6)])
def test_custom_coefficient_modulus(N, bits, k):
    """k primes from SEAL's get_primes rule (data primes of `bits` bits + a special prime one bit larger): covers the
    three NTT arithmetic classes (<2^45 signed-lazy, <2^49 FP64-assisted, larger Shoup) and other limb counts."""
    from oracle.bfv_oracle import get_primes
    data = get_primes(N, bits, k - 1)
    primes = data + [p for p in get_primes(N, bits + 1, 2) if p not in data][:1]
    f, o = make_pair(N, primes)
    try:
        assert f.primes == primes == o.primes and f.k == k
        check_ops(f, o, seed=k)
    finally:
        f.close()


def test_sixteen_primes_n8192():
    """upper end of the 3-16 limb sweep (L = 15 register-resident base conversion)."""
    from oracle.bfv_oracle import get_primes
    data = get_primes(8192, 50, 15)
    primes = data + get_primes(8192, 51, 1)
    f, o = make_pair(8192, primes)
    try:
        check_ops(f, o, seed=16)
    finally:
        f.close()


def test_value_range_and_padding_edges():
    f, o = make_pair(4096)
    try:
        t = f.t
        for data in ([t // 2], [-(t // 2)], [0], [-1, 1], list(range(-5, 6)), [t // 2, -(t // 2), 1]):
            f.set_encrypt_nonce(9)
            ct = f.createCiphertext(data)
            assert np.array_equal(ct.export()[0], o.encrypt_slots(data, 9))
            assert np.array_equal(f.decryptCiphertext(ct), o.expand(data))
        # exactly N values: no padding
        full = np.arange(4096, dtype=np.int64) - 2048
        assert np.array_equal(f.decryptCiphertext(f.createCiphertext(full)), full)
        # products that wrap around t
        big = f.createCiphertext([t // 2, -(t // 2), 1000])
        got = f.decryptCiphertext(big.multiply(big))[:3]
        want = [((v * v + t // 2) % t) - t // 2 for v in (t // 2, -(t // 2), 1000)]
        assert list(got) == want
    finally:
        f.close()


def test_error_behaviour():
    from abc_b200 import AbcError, CudaCiphertextFactory
    f, _ = make_pair(4096)
    g = CudaCiphertextFactory(4096, seed=1)
    try:
        with pytest.raises(AbcError):
            f.createCiphertext([])                      # the reference calls .back() on an empty vector (UB); we throw
        with pytest.raises(AbcError):
            f.createCiphertext(list(range(4097)))       # SealCiphertextFactory.cpp:106-110
        a, b = f.createCiphertext([1, 2]), g.createCiphertext([1, 2])
        with pytest.raises(AbcError):
            a.add(b)                                    # ciphertext of another factory (SealCiphertext.cpp:36-50)
        with pytest.raises(AbcError):
            a.rotateRows(-2048)
        with pytest.raises(AbcError):
            CudaCiphertextFactory(4096, primes=[97, 193])          # not = 1 mod 2N
        with pytest.raises(AbcError):
            CudaCiphertextFactory(5000)                             # not a power of two
        with pytest.raises(AbcError):
            CudaCiphertextFactory(2048)                             # SEAL's default has no special prime below 4096
    finally:
        f.close(); g.close()


def test_transparent_results_are_detected_per_instance():
    """SEAL throws std::logic_error on results whose c1 is zero (x --- x on clones, x *** 0); the C-ABI
    reports them (abc_is_transparent), the C++ drop-in throws when asked to (abc_driver kats, `transparent.*`)."""
    from abc_b200 import CudaCiphertextFactory
    f = CudaCiphertextFactory(4096, seed=SEED, batch=3, galois_steps=[1])
    try:
        x = f.createCiphertext(np.arange(3 * 8).reshape(3, 8))
        assert not x.isTransparent().any()
        assert x.subtract(x.clone()).isTransparent().all()
        assert x.multiplyPlain([0]).isTransparent().all()
        assert not x.add(x).isTransparent().any()
        assert not x.multiply(x).isTransparent().any()
        assert not x.rotateRows(1).isTransparent().any()          # a deferred handle is resolved first
        # one instance only: instance 1 of y equals instance 1 of x, the others differ
        w = x.export()
        w2 = w.copy(); w2[0, 1, 0, 5] ^= np.uint64(1); w2[2, 1, f.L - 1, 4095] ^= np.uint64(1)
        d = f.importCiphertext(w2).subtract(x)
        assert list(d.isTransparent()) == [False, True, False]
    finally:
        f.close()


@pytest.mark.parametrize("limit_mib", ["0", "1", None])
def test_buffer_free_list_limits_do_not_change_results(limit_mib, monkeypatch):
    """Ciphertext buffers are recycled through a per-context free list (lib.cu: salloc / sfree).  With the list disabled
    (ABC_BLOCK_CACHE_MIB=0: every release goes to the driver), with a limit that parks one small block at most, and with
    the default, a program that allocates and drops handles of both buffer sizes gives the oracle's coefficients."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    if limit_mib is not None:
        monkeypatch.setenv("ABC_BLOCK_CACHE_MIB", limit_mib)
    o = Oracle(4096, seed=SEED)
    f = CudaCiphertextFactory(4096, seed=SEED, batch=2)
    try:
        rng = np.random.default_rng(12)
        d = rng.integers(-50, 51, size=(2, 32), dtype=np.int64)
        f.set_encrypt_nonce(3)
        x = f.createCiphertext(d)
        w = [o.encrypt_slots(d[i], 3 * 2 + i) for i in range(2)]
        for rep in range(6):                       # handles of the previous round are dropped while new ones are made
            y = x.multiply(x)
            z = y.rotateRows(3).add(x)
            x = z.subtract(y.clone())
            for i in range(2):
                yw = o.mul_relin(w[i], w[i])
                w[i] = o.sub(o.add(o.rotate_rows(yw, 3), w[i]), yw)
            del y, z
        got = x.export()
        for i in range(2):
            assert np.array_equal(got[i], w[i])
    finally:
        f.close()


def test_key_import_round_trip():
    """abc_key_import: keys produced elsewhere (the oracle here; SEAL's raw key data has the same layout)."""
    from abc_b200 import CudaCiphertextFactory, KEY_GALOIS, KEY_PUBLIC, KEY_RELIN, KEY_SECRET
    from oracle.bfv_oracle import Oracle
    o = Oracle(4096, seed=99)
    f = CudaCiphertextFactory(4096, seed=1, keygen=False)
    try:
        f.import_key(KEY_SECRET, o.secret_key()); f.import_key(KEY_PUBLIC, o.public_key())
        f.import_key(KEY_RELIN, o.relin_key())
        for e in o.galois_elts():
            f.import_key(KEY_GALOIS, o.galois_key(e), e)
        rng = np.random.default_rng(5)
        d = rng.integers(0, 1025, 4096)
        a_w = o.encrypt_slots(d, 3)
        a = f.importCiphertext(a_w[None])
        assert np.array_equal(f.decryptCiphertext(a), d)
        assert np.array_equal(a.multiply(a).export()[0], o.mul_relin(a_w, a_w))
        assert np.array_equal(a.rotateRows(63).export()[0], o.rotate_rows(a_w, 63))
    finally:
        f.close()


def test_n65536_k31_generic_limb_count():
    """BASELINE.json configs[4] parameter set on one GPU: N = 65536, 30 x 55-bit + one 56-bit prime (SEAL has no default
    above 32768), t = 786433; only the keys the deep chain needs (relin + rotate by 1).  Exercises the two-pass NTT with
    A = 3 and the generic (L > 15) base-conversion kernels."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle, get_primes
    N = 65536
    data = get_primes(N, 55, 30)
    primes = data + get_primes(N, 56, 1)
    o = Oracle(N, primes=primes, seed=SEED, galois_steps=[1])
    f = CudaCiphertextFactory(N, primes=primes, seed=SEED, galois_steps=[1])
    try:
        assert f.k == 31 and f.t == o.t == 786433
        rng = np.random.default_rng(65)
        d = rng.integers(0, 4, N)
        f.set_encrypt_nonce(1)
        a = f.createCiphertext(d)
        a_w = o.encrypt_slots(d, 1)
        assert np.array_equal(a.export()[0], a_w), "encrypt"
        sq_w = o.mul_relin(a_w, a_w)
        sq = a.multiply(a)
        assert np.array_equal(sq.export()[0], sq_w), "mul+relin"
        rot = sq.rotateRows(1)
        assert np.array_equal(rot.export()[0], o.rotate_rows(sq_w, 1)), "rotate"
        want = np.concatenate([np.roll((d * d)[:N // 2], -1), np.roll((d * d)[N // 2:], -1)])
        assert np.array_equal(f.decryptCiphertext(rot), want)
    finally:
        f.close()


def test_limb_sharded_key_switch_two_gpus():
    """BASELINE.json configs[4] mechanism at test size: RNS limbs sharded over 2 GPUs, NCCL all-gather in front of the
    key-switch ModUp, owned limbs bit-exact against the oracle (tools/shard_check.py).  Needs 2 visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run by hand: gpurun --gpus 2 -- torchrun ... tools/shard_check.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "shard_check.py"), "8192"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "shard_check ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


KS_VARIANTS = {
    "one_launch": {"ABC_KS_ONE_LAUNCH": "1"},                                   # ksfused.cu: accumulators in shared memory
    "chained": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "1", "ABC_KS_SPLIT_MAXB": "0"},   # kschain.cu: ModUp + tail rows in one grid
    "accumulating": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "1", "ABC_KS_RED": "1"},  # ksred.cu: ModUp rows bulk-reduce into the accumulators
    "split_rows": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "1", "ABC_KS_SPLIT_MAXB": "8", "ABC_KS_SPLIT_FORCE": "1"},  # ks14.cu: rows of half a limb (N = 8192 at batch <= 8)
    "two_launch": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "0"},              # ModUp launch, tail launch (T as images)
    "two_launch_canonical_T": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "0", "ABC_KS_NO_IMAGE": "1"},
    "unfused_tail": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_UNFUSED": "1"},          # separate inner-product kernel
}


@pytest.mark.parametrize("N", [4096, 8192])
@pytest.mark.parametrize("variant", sorted(KS_VARIANTS))
def test_key_switch_variants_agree(N, variant, monkeypatch):
    """The exact-double key switch exists in several launch structures (the default is picked by size); each is forced
    here and must give the oracle's coefficients for relinearisation, single and NAF rotations, the fused rotate+add
    and a batch of 3 (more instances than the chained schedule's skew would be covered by the bench's result check)."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    for kk, vv in KS_VARIANTS[variant].items():
        monkeypatch.setenv(kk, vv)
    o = Oracle(N, seed=SEED)
    f = CudaCiphertextFactory(N, seed=SEED, batch=3)
    try:
        rng = np.random.default_rng(N + len(variant))
        da, db = rng.integers(0, 1025, (3, N)), rng.integers(0, 1025, (3, N))
        a_w = np.stack([o.encrypt_slots(da[i], 10 + i) for i in range(3)])
        b_w = np.stack([o.encrypt_slots(db[i], 20 + i) for i in range(3)])
        a, b = f.importCiphertext(a_w), f.importCiphertext(b_w)
        prod = a.multiply(b).export()
        for i in range(3):
            assert np.array_equal(prod[i], o.mul_relin(a_w[i], b_w[i])), "mul+relin inst %d" % i
        for steps in (1, -24, 63):
            rot = a.rotateRows(steps).export()
            for i in range(3):
                assert np.array_equal(rot[i], o.rotate_rows(a_w[i], steps)), "rotate %d inst %d" % (steps, i)
        s = b.add(a.rotateRows(4)).export()          # deferred rotation consumed by the add: addend in the ModDown
        for i in range(3):
            assert np.array_equal(s[i], o.add(b_w[i], o.rotate_rows(a_w[i], 4))), "rotate+add inst %d" % i
    finally:
        f.close()


def test_chained_key_switch_large_batch_matches_two_launch(monkeypatch):
    """48 instances (3x the chained schedule's skew, so ModUp rows of later instances are in flight while earlier ones
    finish): rotate, rotate+add and mul+relin coefficients of the chained grid equal the two-launch sequence's, and
    instance 0 / 47 equal the oracle's."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    N, B = 8192, 48
    rng = np.random.default_rng(7)
    da, db = rng.integers(0, 1025, (B, N)), rng.integers(0, 1025, (B, N))
    outs = {}
    for variant in ("chained", "two_launch"):
        for kk, vv in KS_VARIANTS[variant].items():
            monkeypatch.setenv(kk, vv)
        f = CudaCiphertextFactory(N, seed=SEED, batch=B)
        try:
            f.set_encrypt_nonce(5)
            a, b = f.createCiphertext(da), f.createCiphertext(db)
            outs[variant] = (a.export(), a.rotateRows(-24).export(), b.add(a.rotateRows(1)).export(), a.multiply(b).export())
        finally:
            f.close()
    for x, y in zip(outs["chained"], outs["two_launch"]):
        assert np.array_equal(x, y)
    o = Oracle(N, seed=SEED)
    a_in, rot = outs["chained"][0], outs["chained"][1]
    for i in (0, B - 1):
        assert np.array_equal(rot[i], o.rotate_rows(a_in[i], -24))


def test_bench_batch_identical_instances_agree():
    """Size-independent property at the bench's full batch (592 instances, N = 8192): the same ciphertext pair imported
    into every instance must give bit-identical coefficients in every instance after mul+relin, a NAF rotation and the
    fused rotate+add (17 760 rows per key-switch grid, every dependency of the chained schedule exercised), and instance 0
    must equal the oracle."""
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    N, B = 8192, 592
    o = Oracle(N, seed=SEED)
    rng = np.random.default_rng(11)
    a_w, b_w = o.encrypt_slots(rng.integers(0, 1025, N), 1), o.encrypt_slots(rng.integers(0, 1025, N), 2)
    f = CudaCiphertextFactory(N, seed=SEED, batch=B)
    try:
        a = f.importCiphertext(np.broadcast_to(a_w, (B,) + a_w.shape))
        b = f.importCiphertext(np.broadcast_to(b_w, (B,) + b_w.shape))
        m_w = o.mul_relin(a_w, b_w)
        r_w = o.rotate_rows(m_w, -24)
        s_w = o.add(b_w, o.rotate_rows(r_w, 4))
        m = a.multiply(b)
        r = m.rotateRows(-24)
        s = b.add(r.rotateRows(4))
        for name, ct, want in (("mul+relin", m, m_w), ("rotate(-24)", r, r_w), ("rotate(4)+add", s, s_w)):
            got = ct.export()
            assert np.array_equal(got[0], want), name + ": instance 0 vs oracle"
            assert (got == got[0]).all(), name + ": instances differ"
        assert np.array_equal(f.decryptCiphertext(s)[B - 1], o.decrypt_slots(s_w))
    finally:
        f.close()


def test_key_switch_grids_survive_gpu_time_slicing():
    """Three processes share the GPU (their contexts are time-sliced, resident CTAs are preempted in the middle of their
    flag waits) while each loops over the chained and the half-limb-row key switch: every result equals the first one and
    the oracle's, no sticky fault (tools/timeslice_check.py).  The evidence asked for where compute-sanitizer is not
    available: the dependency-ordered grids do not rest on exclusive use of the device."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "timeslice_check.py"), "--procs", "3", "--seconds", "6"],
                       capture_output=True, text=True, timeout=600)
    if p.returncode == 0 and "timeslice_check skipped" in p.stdout:
        pytest.skip("the device does not admit several processes (exclusive compute mode)")
    assert p.returncode == 0 and "timeslice_check ok" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
