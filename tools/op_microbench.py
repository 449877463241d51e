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
    """Returns (mean, p50, p99) ms of one batched launch sequence, each repetition timed between two synchronisations."""
    fn()  # warm
    f.sync()
    ts = []
    for _ in range(reps):
        if flush:
            f.flush_l2(256 << 20)
        f.timer_start()
        fn()
        ts.append(f.timer_stop())
    ts.sort()
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[min(len(ts) - 1, int(0.99 * len(ts)))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="8192,16384")
    ap.add_argument("--batches", default="1,16,256")
    ap.add_argument("--limbs", default="0", help="comma list of k (3..16); 0 = BFVDefault(N)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--max-gib", type=float, default=40.0, help="skip configurations whose key-switch scratch alone exceeds this")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    rows = []
    peaks = {}
    for N in [int(s) for s in args.sizes.split(",")]:
      for klimbs in [int(v) for v in args.limbs.split(",")]:
        primes = None
        if klimbs:
            data = seal_primes(N, 50, klimbs - 1)
            primes = data + seal_primes(N, 51, 1, skip=data)
        for B in [int(b) for b in args.batches.split(",")]:
            kk = klimbs if klimbs else {4096: 3, 8192: 5, 16384: 9, 32768: 16}[N]
            if B * kk * (kk - 1) * N * 8 / 2**30 > args.max_gib:
                print("N=%5d k=%2d B=%4d skipped (ModUp scratch above %.0f GiB)" % (N, kk, B, args.max_gib), flush=True)
                continue
            f = CudaCiphertextFactory(N, primes=primes, batch=B, galois_steps=[1, 4, 8, -32, 64, -1])
            L, k = f.L, f.k
            arq = f.ntt_arith_class()
            if arq not in peaks:
                peaks[arq] = f.measure_butterfly_peak(arq)
            if 0 not in peaks:
                peaks[0] = f.measure_butterfly_peak(0)
            logn = N.bit_length() - 1
            bf_row = (N // 2) * logn
            ks_rows = k * L + 2 * k                      # ModUp + forward rows, inverse rows of one key switch
            qbits = sum(int(p).bit_length() for p in f.primes[:L])
            W2 = L + (32 + int(f.t).bit_length() + qbits + 8 + 42) // 43          # FP64 BEHZ rows per polynomial (lib.cu)
            f64_behz = arq == 3 and L <= 8
            behz_ideal_s = (7 * W2 * bf_row / peaks[3]) if f64_behz else (7 * L * bf_row / peaks[arq] + 7 * (L + 1) * bf_row / peaks[0])
            rng = np.random.default_rng(1)
            a = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
            b = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
            out = f.allocCiphertext()
            lib = f._lib
            flush = B * 2 * L * N * 8 * 3 < (200 << 20)
            ops = {"add": (lambda: f._ck(lib.abc_add(f._h, out._h, a._h, b._h)), 48 * L * N, 0, 0.0),
                   "mul_relin": (lambda: f._ck(lib.abc_mul_relin(f._h, out._h, a._h, b._h)), 16 * L * N * (L + 4), 1,
                                 behz_ideal_s + ks_rows * bf_row / peaks[arq])}
            for ks_ in ROT_STEPS:
                nks = 1 if ks_ in (1, 4) else 2
                ops["rotate(%d)" % ks_] = ((lambda ks_=ks_: f._ck(lib.abc_rotate_rows(f._h, out._h, a._h, ks_))),
                                           16 * L * N * (L + 3) * nks, nks, nks * ks_rows * bf_row / peaks[arq])
            # add(rotate_rows(a, 1), b) as one key switch (what a rotate-and-sum ladder step costs)
            ops["rotate(1)+add"] = ((lambda: f._ck(lib.abc_rotate_rows_add(f._h, out._h, a._h, 1, b._h))),
                                    16 * L * N * (L + 3) + 16 * L * N, 1, ks_rows * bf_row / peaks[arq])
            for name, (fn, alg_bytes, nks, ideal_s) in ops.items():
                ms, p50, p99 = time_op(f, fn, args.reps, flush)
                # keys are shared by the batch: amortise the key term
                key_bytes = 16 * L * (L + 1) * N * nks
                bytes_batch = B * (alg_bytes - key_bytes) + key_bytes
                rows.append({"N": N, "k": f.k, "batch": B, "op": name, "arith_class": arq, "ms": ms, "p50_ms": p50, "p99_ms": p99,
                             "ops_per_s": B / (ms * 1e-3), "latency_us_per_launch_sequence": p50 * 1e3,
                             "hbm_gbs": bytes_batch / (ms * 1e-3) / 1e9, "hbm_frac": bytes_batch / (ms * 1e-3) / 1e9 / hbm,
                             # butterflies of the op's transforms at the register-resident peak of their arithmetic class / time
                             "int_roofline_frac": (B * ideal_s) / (ms * 1e-3) if ideal_s else None})
                print("N=%5d k=%2d B=%4d %-13s %9.3f ms (p99 %8.3f) %11.0f ops/s  HBM %5.0f GB/s (%4.1f%%)  butterfly-roof %s" % (
                    N, f.k, B, name, ms, p99, B / (ms * 1e-3), rows[-1]["hbm_gbs"], 100 * rows[-1]["hbm_frac"],
                    "%.2f" % rows[-1]["int_roofline_frac"] if ideal_s else "-"), flush=True)
            del a, b, out
            f.close()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
