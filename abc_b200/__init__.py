"""abc_b200 — B200 (sm_100a) BFV ciphertext backend for MarbleHE/ABC.

Product path: abc_b200/csrc (CUDA kernels + C ABI, built to abc_b200/lib/libabc_b200.so),
abc_b200/cpp (C++ CudaCiphertextFactory / CudaCiphertext behind ABC's AbstractCiphertext* interfaces)
and this Python binding of the C ABI.  There is no CPU fallback.
"""
from .context import (AbcError, CudaCiphertext, CudaCiphertextFactory, CudaPlaintext, KEY_GALOIS, KEY_PUBLIC,
                      KEY_RELIN, KEY_SECRET, seal_parameters_from_bytes)

__all__ = ["AbcError", "CudaCiphertext", "CudaCiphertextFactory", "CudaPlaintext", "KEY_SECRET", "KEY_PUBLIC",
           "KEY_RELIN", "KEY_GALOIS", "seal_parameters_from_bytes"]
