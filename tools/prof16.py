import sys; sys.path.insert(0,".")
import numpy as np
from abc_b200 import CudaCiphertextFactory
B=256
f=CudaCiphertextFactory(16384,batch=B,seed=1,galois_steps=[1])
rng=np.random.default_rng(0)
a=f.createCiphertext(rng.integers(0,1025,(B,16384))); b=f.createCiphertext(rng.integers(0,1025,(B,16384)))
a.multiply(b); f.sync(); f.profile_enable(True); a.multiply(b); r=a.rotateRows(1); f.sync()
for k in f.profile(): print(k)
