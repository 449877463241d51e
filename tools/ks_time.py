"""A/B timing of the key-switch ops for library variants (tools/build_variant.sh; select with ABC_B200_LIB).
Prints one JSON line: ms per rotate / rotate+add / mul+relin at (N, batch), and a digest of each result so that
variants that must be bit-identical can be compared without the oracle.

  python tools/ks_time.py [--n 8192] [--batch 592] [--reps 20] [--tag name]
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

os.environ.setdefault("ABC_EAGER_ROTATE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tag", default=os.environ.get("ABC_B200_LIB", "default"))
    ap.add_argument("--ops", default="rotate,rotate_add,mul_relin")
    args = ap.parse_args()
    N, B = args.n, args.batch
    f = CudaCiphertextFactory(N, batch=B, seed=4673838)
    rng = np.random.default_rng(1)
    f.set_encrypt_nonce(1)
    a = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
    b = f.createCiphertext(rng.integers(0, 1025, size=(B, N) if B > 1 else N))
    out = f.allocCiphertext()
    lib = f._lib
    ops = {"rotate": lambda: f._ck(lib.abc_rotate_rows(f._h, out._h, a._h, 1)),
           "rotate_add": lambda: f._ck(lib.abc_rotate_rows_add(f._h, out._h, a._h, 1, b._h)),
           "mul_relin": lambda: f._ck(lib.abc_mul_relin(f._h, out._h, a._h, b._h)),
           "add": lambda: f._ck(lib.abc_add(f._h, out._h, a._h, b._h))}
    res = {"tag": args.tag, "N": N, "batch": B}
    for name in args.ops.split(","):
        fn = ops[name]
        fn()
        f.sync()
        w = out.export()
        res[name + "_digest"] = hashlib.sha1(np.ascontiguousarray(w[: min(B, 4)]).tobytes()).hexdigest()[:12]
        for _ in range(3):
            fn()
        f.sync()
        f.timer_start()
        for _ in range(args.reps):
            fn()
        res[name + "_ms"] = round(f.timer_stop() / args.reps, 4)
        res[name + "_kops"] = round(B / res[name + "_ms"], 1)
    print(json.dumps(res), flush=True)
    f.close()


if __name__ == "__main__":
    main()
