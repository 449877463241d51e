"""Dependency-ordered key-switch grids under GPU time slicing: several processes share ONE GPU (the driver time-slices
their contexts, preempting resident CTAs in the middle of their flag waits), each runs the chained / half-limb-row key
switch in a loop and compares every result with the first one and with the oracle.  A wait that gave up would raise the
sticky fault (abc_faulted) and corrupt the digest.

  python tools/timeslice_check.py [--procs 3] [--seconds 8]      (parent: spawns the workers on cuda:0)
"""
import argparse
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(idx, seconds):
    os.environ["ABC_EAGER_ROTATE"] = "1"
    from abc_b200 import CudaCiphertextFactory
    from oracle.bfv_oracle import Oracle
    # worker 0: N = 8192, batch 64 (chained grid, several waves); worker 1: N = 16384 batch 8 (half-limb rows, partner
    # exchange); others: N = 8192 batch 4 (half-limb rows at small batch)
    N, B = [(8192, 64), (16384, 8), (8192, 4)][min(idx, 2)]
    seed = 4673838
    try:
        f = CudaCiphertextFactory(N, batch=B, seed=seed, galois_steps=[1, 4])
    except Exception as e:   # e.g. the device is in an exclusive compute mode: a second process gets no context
        print("worker %d skipped: no context on the shared device (%s)" % (idx, e), flush=True)
        sys.exit(77)
    o = Oracle(N, seed=seed, galois_steps=[1, 4])
    rng = np.random.default_rng(idx)
    d = rng.integers(0, 1025, size=(B, 64), dtype=np.int64)
    f.set_encrypt_nonce(5)
    x = f.createCiphertext(d)
    w0 = o.encrypt_slots(d[0], 5 * B)
    want = o.add(o.rotate_rows(o.mul_relin(w0, w0), 1), o.rotate_rows(w0, 4))
    first, n = None, 0
    t_end = time.time() + seconds
    while time.time() < t_end:
        for _ in range(10):
            y = x.multiply(x).rotateRows(1).add(x.rotateRows(4))
        got = y.export()
        h = hashlib.sha256(got.tobytes()).hexdigest()
        if first is None:
            first = h
            assert np.array_equal(got[0], want), "worker %d: result differs from the oracle" % idx
        assert h == first, "worker %d: iteration %d differs from the first" % (idx, n)
        n += 10
    assert f._lib.abc_faulted(f._h) == 0, "worker %d: sticky fault raised" % idx
    f.close()
    print("worker %d ok: N=%d batch=%d, %d programs, digest %s" % (idx, N, B, n, first[:12]), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--worker", type=int, default=-1)
    args = ap.parse_args()
    if args.worker >= 0:
        return worker(args.worker, args.seconds)
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(i), "--seconds", str(args.seconds)],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for i in range(args.procs)]
    ok, ran = True, 0
    for p in ps:
        out, _ = p.communicate(timeout=600)
        print(out.strip()[-600:])
        if p.returncode == 77:
            continue
        ran += 1
        ok = ok and p.returncode == 0 and "ok:" in out
    if ran < 2:   # nothing was shared: the device does not admit several processes
        print("timeslice_check skipped: %d of %d processes got a context on cuda:0" % (ran, args.procs))
        sys.exit(0 if ok else 1)
    print("timeslice_check %s: %d processes sharing cuda:0" % ("ok" if ok else "FAILED", ran))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
