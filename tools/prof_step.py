"""One small invocation of the bench workload for ncu (no torch, no CPU baseline).
usage: python tools/prof_step.py [batch] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertextFactory  # noqa: E402
from bench import N_POLY, N_VEC, SEED, program_gpu, synth_inputs, expected_slot0  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f = CudaCiphertextFactory(N_POLY, batch=B, seed=SEED)
xs, ys = synth_inputs(B, 0)
x, y = f.createCiphertext(xs), f.createCiphertext(ys)
for _ in range(steps):
    s = program_gpu(x, y)
got = np.atleast_2d(f.decryptCiphertext(s))[:, 0]
want = expected_slot0(xs, ys) % f.t
want = np.where(want > f.t // 2, want - f.t, want)
assert np.array_equal(got, want)
print("ok", B, steps, f.launch_count())
f.close()
