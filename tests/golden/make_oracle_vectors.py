"""Coefficient-level regression digests of the ORACLE itself (not of SEAL: the reference holds no ciphertext fixtures).
They pin oracle/bfv_oracle.c — key generation and encryption under the ChaCha20 sampler, BEHZ multiply, relinearisation,
Galois rotations with NAF chains, plain ops, decryption — against accidental change: tests/test_oracle.py recomputes them.

  python tests/golden/make_oracle_vectors.py        (rewrites tests/golden/oracle_vectors.json)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bfv_oracle import Oracle  # noqa: E402

SEED = 4673838


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()[:32]


def vectors(N):
    o = Oracle(N, seed=SEED, galois_steps=[1, 4, -32, 64])
    x = o.encrypt_slots([3, 3, 1, 4, 5, 9], 7)
    y = o.encrypt_slots([0, 1, 2, 1, 10, 21], 8)
    plain = o.encode(o.expand([2, -3, 5]))
    prod = o.mul_relin(x, y)
    out = {
        "secret_key": digest(o.secret_key()), "public_key": digest(o.public_key()), "relin_key": digest(o.relin_key()),
        "galois_key_step1": digest(o.galois_key(o.elt_from_step(1))),
        "encrypt_nonce7": digest(x), "add": digest(o.add(x, y)), "sub": digest(o.sub(x, y)), "negate": digest(o.negate(x)),
        "mul_relin": digest(prod), "square": digest(o.mul_relin(x, x)),
        "rotate_1": digest(o.rotate_rows(x, 1)), "rotate_-28": digest(o.rotate_rows(x, -28)), "rotate_69": digest(o.rotate_rows(prod, 69)),
        "add_plain": digest(o.add_plain(x, plain)), "sub_plain": digest(o.sub_plain(x, plain)),
        "multiply_plain": digest(o.multiply_plain(x, plain)),
        "decrypt_of_product": [int(v) for v in o.decrypt_slots(prod)[:6]],
        "noise_budget_of_product": int(o.noise_budget(prod)),
    }
    return out


def main():
    res = {"_comment": "sha256 (first 32 hex digits) of little-endian u64 words; seed %d; made by tests/golden/make_oracle_vectors.py" % SEED,
           "4096": vectors(4096), "8192": vectors(8192)}
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json"), "w"), indent=1)
    print(json.dumps(res["4096"], indent=1))


if __name__ == "__main__":
    main()
