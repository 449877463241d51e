"""NTT kernel microbenchmark: python tools/bench_ntt.py [N] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 40
f = CudaCiphertextFactory(N, keygen=False)
logn = N.bit_length() - 1
peaks = [f.measure_butterfly_peak(a) for a in (0, 1, 2, 3)]
print('butterfly microbench peaks (Shoup, FP, FP-lazy, pure-FP64): %s G/s' % ['%.0f' % (p / 1e9) for p in peaks])
peak = peaks[0]
for name, mi in (("q0", 0), ("bsk0", f.k)):
    for inv in (False, True):
        ms = f.bench_ntt(mi, rows, 10, inv) / 10
        bf = rows * (N // 2) * logn / (ms * 1e-3)
        print("N=%d %s %s: %.1f us/launch  %.2f Mrows/s  %.1f Gbf/s (%.0f%% of butterfly microbench peak %.0f G/s)  %.0f GB/s" % (
            N, name, "inv" if inv else "fwd", ms * 1e3, rows / ms / 1e3, bf / 1e9, 100 * bf / peak, peak / 1e9, rows * N * 16 / ms / 1e6))
f.close()
