"""Where the end-to-end step of bench.py spends its device time: per-kernel CUDA-event times of one e2e step
(createCiphertext x2 from pinned host slots, program, decryptCiphertext to host) next to the step's total.
usage: python tools/e2e_breakdown.py [batch]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abc_b200 import CudaCiphertext, CudaCiphertextFactory  # noqa: E402
from bench import N_POLY, N_VEC, SEED, program_gpu, synth_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
f = CudaCiphertextFactory(N_POLY, batch=B, seed=SEED)
xs, ys = synth_inputs(B, 0)
hx, hy = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
hout = torch.empty((B, N_POLY), dtype=torch.int64).pin_memory()
lib = f._lib


def step():
    a, b = C.c_void_p(), C.c_void_p()
    f._ck(lib.abc_encode_encrypt(f._h, hx.data_ptr(), N_VEC, 0, C.byref(a)))
    f._ck(lib.abc_encode_encrypt(f._h, hy.data_ptr(), N_VEC, 0, C.byref(b)))
    r = program_gpu(CudaCiphertext(f, a), CudaCiphertext(f, b))
    f._ck(lib.abc_decrypt_decode(f._h, r._h, hout.data_ptr()))


for _ in range(3):
    step()
f.timer_start()
for _ in range(5):
    step()
tot = f.timer_stop() / 5
f.profile_enable(True)
step()
prof = f.profile()
f.profile_enable(False)
ks = sum(r["ms"] for r in prof)
print("e2e step %.3f ms; kernels %.3f ms; copies + gaps %.3f ms" % (tot, ks, tot - ks))
prog = {"sub", "add"}
for r in sorted(prof, key=lambda r: -r["ms"]):
    print("  %-26s x%-3d %8.4f ms" % (r["kernel"], r["launches"], r["ms"]))
f.close()
