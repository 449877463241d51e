"""A few batched ops at N = 16384 (the reference's default factory size) for ncu: python tools/prof_n16384.py [batch]"""
import os
import sys

import numpy as np

os.environ.setdefault("ABC_EAGER_ROTATE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
f = CudaCiphertextFactory(16384, batch=B, seed=4673838, galois_steps=[1])
rng = np.random.default_rng(0)
a = f.createCiphertext(rng.integers(0, 1025, (B, 16384)))
b = f.createCiphertext(rng.integers(0, 1025, (B, 16384)))
for _ in range(2):
    r = a.rotateRows(1)
    m = a.multiply(b)
f.sync()
print("ok", B, f.launch_count())
f.close()
