"""Extreme-residue parity of the transforms and of the key switch on every arithmetic class (parity gap i).

The range plans of the lazy classes (csrc/ntt.cuh: "|x| < 2^45 / |v| <= 0.6 q", the exact-double inner product and ModDown of
csrc/limb.cuh) are stressed by inputs a valid encryption never produces: every residue q - 1, alternating 0 / q - 1, a
single spike — in the ciphertext AND in the key-switching key (imported through abc_key_import).  The oracle's
switch_key (oracle/bfv_oracle.c, SEAL 3.6.5 Evaluator::switch_key_inplace) takes the same raw key, so the comparison is
coefficient-exact.  ABC_FORCE_AR selects a lower class for the same primes (0 Shoup, 1 FP64-assisted, 2 signed-lazy).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 4673838
CLASSES = {"shoup": "0", "fp": "1", "fp_lazy": "2", "f64": None}


def _patterns(q, N, rng):
    q1 = np.uint64(q - 1)
    rows = {"max": np.full(N, q1, dtype=np.uint64), "zero": np.zeros(N, dtype=np.uint64)}
    alt = np.zeros(N, dtype=np.uint64); alt[::2] = q1
    rows["alt"] = alt
    alt2 = np.zeros(N, dtype=np.uint64); alt2[1::2] = q1
    rows["alt_odd"] = alt2
    sp = np.zeros(N, dtype=np.uint64); sp[N - 1] = q1
    rows["spike_last"] = sp
    sp0 = np.zeros(N, dtype=np.uint64); sp0[0] = np.uint64(1)
    rows["spike_first"] = sp0
    half = np.full(N, np.uint64(q // 2), dtype=np.uint64); half[N // 2:] = np.uint64(q // 2 + 1)
    rows["half"] = half
    rows["rand"] = rng.integers(0, q, size=N, dtype=np.uint64)
    return rows


@pytest.mark.parametrize("N,cls", [(n, c) for n in (4096, 8192) for c in sorted(CLASSES)] + [(16384, "f64"), (16384, "fp")])
def test_probe_ntt_extreme_rows_every_class(N, cls, monkeypatch):
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    if CLASSES[cls] is not None:
        monkeypatch.setenv("ABC_FORCE_AR", CLASSES[cls])
    o = Oracle(N, seed=SEED)
    f = CudaCiphertextFactory(N, seed=SEED, keygen=False)
    try:
        rng = np.random.default_rng(N)
        for mi in range(f.k):
            pats = _patterns(f.primes[mi], N, rng)
            rows = np.stack(list(pats.values()))
            got = f.probe_ntt(mi, rows)
            for r, name in enumerate(pats):
                assert np.array_equal(got[r], o.ntt_fwd(mi, rows[r])), "forward %s mod %d class %s" % (name, mi, cls)
            # the forward outputs of extreme rows are themselves extreme inverse inputs (all-equal, alternating sign ...)
            assert np.array_equal(f.probe_ntt(mi, got, inverse=True), rows), "inverse round trip mod %d class %s" % (mi, cls)
            inv = f.probe_ntt(mi, rows, inverse=True)
            for r, name in enumerate(pats):
                assert np.array_equal(inv[r], o.ntt_inv(mi, rows[r])), "inverse %s mod %d class %s" % (name, mi, cls)
    finally:
        f.close()


def _galois_coeff(poly, elt, q):
    """GaloisTool::apply_galois on one limb in coefficient form: out[i * elt mod N] = +-in[i] (sign from bit log2 N)."""
    N = poly.shape[-1]
    i = np.arange(N, dtype=np.int64)
    j = (i * elt) % (2 * N)
    out = np.zeros_like(poly)
    neg = j >= N
    v = poly.copy()
    v[neg] = (np.uint64(q) - v[neg]) % np.uint64(q)
    out[j % N] = v
    return out


KS_PATHS = {"default": {}, "chained": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_SPLIT_MAXB": "0"}, "two_launch": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_CHAIN": "0"},
            "one_launch": {"ABC_KS_ONE_LAUNCH": "1"}, "accumulating": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_RED": "1"},
            "persistent": {"ABC_KS_ONE_LAUNCH": "0", "ABC_KS_SPLIT_MAXB": "0", "ABC_KS_PERSIST": "1"}}


@pytest.mark.parametrize("N,cls,path", [(n, "f64", p) for n in (4096, 8192) for p in sorted(KS_PATHS)] +
                         [(n, c, "default") for n in (4096, 8192) for c in ("shoup", "fp", "fp_lazy")] +
                         [(16384, "f64", "default"), (16384, "fp", "default")])   # 16384: 48/49-bit primes, the wide exact-double plan
def test_key_switch_extreme_residues_and_keys(N, cls, path, monkeypatch):
    """rotate_rows with an imported Galois key whose every residue is q_I - 1 (then alternating, then random) applied to
    ciphertexts of extreme residues; expected value = (sigma(c0), 0) + switch_key(sigma(c1), key) from the oracle."""
    from abc_b200 import CudaCiphertextFactory, KEY_GALOIS
    from oracle.bfv_oracle import Oracle
    if CLASSES[cls] is not None:
        monkeypatch.setenv("ABC_FORCE_AR", CLASSES[cls])
    for kk, vv in KS_PATHS[path].items():
        monkeypatch.setenv(kk, vv)
    monkeypatch.setenv("ABC_EAGER_ROTATE", "1")
    o = Oracle(N, seed=SEED)
    f = CudaCiphertextFactory(N, seed=SEED, batch=2, galois_steps=[1])
    try:
        L, k = f.L, f.k
        elt = f.elt_from_step(1)
        rng = np.random.default_rng(N + 17)
        qk = np.array(f.primes, dtype=np.uint64)
        keys = {}
        kmax = np.zeros((L, 2, k, N), dtype=np.uint64)
        kmax[:] = (qk - np.uint64(1))[None, None, :, None]
        keys["max"] = kmax
        kalt = kmax.copy(); kalt[..., 1::2] = 0
        keys["alt"] = kalt
        keys["rand"] = np.stack([rng.integers(0, int(qk[i]), size=(L, 2, N), dtype=np.uint64) for i in range(k)], axis=2)
        for kname, key in keys.items():
            f.import_key(KEY_GALOIS, key, elt)
            for pa, pb in (("max", "alt"), ("alt_odd", "spike_last"), ("half", "rand"), ("spike_first", "max")):
                cts = []
                for pname in (pa, pb):
                    ct = np.zeros((2, L, N), dtype=np.uint64)
                    for i in range(L):
                        pats = _patterns(f.primes[i], N, rng)
                        ct[0, i], ct[1, i] = pats[pname], pats[pname][::-1]
                    cts.append(ct)
                a_w = np.stack(cts)                      # two instances with different patterns
                a = f.importCiphertext(a_w)
                got = a.rotateRows(1).export()
                for inst in range(2):
                    s0 = np.stack([_galois_coeff(a_w[inst, 0, i], elt, f.primes[i]) for i in range(L)])
                    s1 = np.stack([_galois_coeff(a_w[inst, 1, i], elt, f.primes[i]) for i in range(L)])
                    want = o.switch_key(np.stack([s0, np.zeros_like(s0)]), s1, key)
                    assert np.array_equal(got[inst], want), "key %s ct %s class %s path %s" % (kname, (pa, pb)[inst], cls, path)
    finally:
        f.close()
