import time, sys, os
sys.path.insert(0, "/root/repo")
t0 = time.time()
from abc_b200 import CudaCiphertextFactory
print("import %.3f" % (time.time() - t0))
for N in (8192, 16384, 16384, 8192, 4096, 32768):
    t0 = time.time()
    f = CudaCiphertextFactory(N, keygen=False)
    f.sync(); t1 = time.time()
    f.keygen(None); f.sync(); t2 = time.time()
    x = f.createCiphertext([1, 2, 3]); y = x.multiply(x); y.rotateRowsInplace(1); f.sync(); t3 = time.time()
    print("N=%d ctx %.3f s keygen %.3f s first ops %.3f s" % (N, t1 - t0, t2 - t1, t3 - t2), flush=True)
    del x, y
    f.close()
