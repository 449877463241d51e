"""Op microbenchmark (BASELINE.json configs[2]): add, mul+relin and rotateRows throughput per (N, k, batch) on one B200,
device-resident operands, CUDA events on the library stream, L2 flushed between timed ops when the working set is small.

  python tools/op_microbench.py [--sizes 8192,16384] [--batches 1,16,256] [--limbs 0] [--out profiles/x.json]
`--limbs k` (3..16) replaces SEAL's default coefficient modulus by k primes from SEAL's get_primes rule
(50-bit data primes + one 51-bit special prime); 0 = BFVDefault(N).
Roofline columns: HBM = algorithmic bytes (add 48LN, mul+relin 16LN(L+4), rotate 16LN(L+3) per key switch,
keys amortised over the batch) / time vs MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import numpy as np

os.environ["ABC_EAGER_ROTATE"] = "1"   # time rotate_rows itself: by default its last key switch is deferred to the consumer
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402

ROT_STEPS = (1, 4, -24, 63)   # NAF weights 1, 1, 2, 2


def seal_primes(N, bits, count, skip=()):
    """SEAL util::get_primes: scan down from 2^bits - 2N + 1 in steps of 2N."""
    def is_prime(n):
        if n < 2:
            return False
        for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if n % p == 0:
                return n == p
        d, s = n - 1, 0
        while d % 2 == 0:
            d //= 2; s += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(s - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True
    out, v = [], (1 << bits) - 2 * N + 1
    while len(out) < count and v > (1 << (bits - 1)):
        if is_prime(v) and v not in skip:
            out.append(v)
        v -= 2 * N
    return out


def time_op(f, fn, reps, flush):
    fn()  # warm
    f.sync()
    total = 0.0
    for _ in range(reps):
        if flush:
            f.flush_l2(256 << 20)
        f.timer_start()
        fn()
        total += f.timer_stop()
    return total / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="8192,16384")
    ap.add_argument("--batches", default="1,16,256")
    ap.add_argument("--limbs", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    rows = []
    for N in [int(s) for s in args.sizes.split(",")]:
        primes = None
        if args.limbs:
            data = seal_primes(N, 50, args.limbs - 1)
            primes = data + seal_primes(N, 51, 1, skip=data)
        for B in [int(b) for b in args.batches.split(",")]:
            f = CudaCiphertextFactory(N, primes=primes, batch=B)
            L = f.L
            rng = np.random.default_rng(1)
            a = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
            b = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
            out = f.allocCiphertext()
            lib = f._lib
            flush = B * 2 * L * N * 8 * 3 < (200 << 20)
            ops = {"add": (lambda: f._ck(lib.abc_add(f._h, out._h, a._h, b._h)), 48 * L * N, 1),
                   "mul_relin": (lambda: f._ck(lib.abc_mul_relin(f._h, out._h, a._h, b._h)), 16 * L * N * (L + 4), 1)}
            for k in ROT_STEPS:
                nks = 1 if k in (1, 4) else 2
                ops["rotate(%d)" % k] = ((lambda k=k: f._ck(lib.abc_rotate_rows(f._h, out._h, a._h, k))),
                                         16 * L * N * (L + 3) * nks, nks)
            # add(rotate_rows(a, 1), b) as one key switch (what a rotate-and-sum ladder step costs)
            ops["rotate(1)+add"] = ((lambda: f._ck(lib.abc_rotate_rows_add(f._h, out._h, a._h, 1, b._h))),
                                    16 * L * N * (L + 3) + 16 * L * N, 1)
            for name, (fn, alg_bytes, nks) in ops.items():
                ms = time_op(f, fn, args.reps, flush)
                # keys are shared by the batch: amortise the key term
                key_bytes = 16 * L * (L + 1) * N * (0 if name == "add" else nks)
                bytes_batch = B * (alg_bytes - key_bytes) + key_bytes
                rows.append({"N": N, "k": f.k, "batch": B, "op": name, "ms": ms, "ops_per_s": B / (ms * 1e-3),
                             "latency_us_per_launch_sequence": ms * 1e3,
                             "hbm_gbs": bytes_batch / (ms * 1e-3) / 1e9, "hbm_frac": bytes_batch / (ms * 1e-3) / 1e9 / hbm})
                print("N=%5d k=%2d B=%4d %-12s %9.3f ms  %12.0f ops/s  HBM %.0f GB/s (%.1f%% of measured %.0f)" % (
                    N, f.k, B, name, ms, B / (ms * 1e-3), rows[-1]["hbm_gbs"], 100 * rows[-1]["hbm_frac"], hbm), flush=True)
            del a, b, out
            f.close()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
