"""CPU tests: the C-ABI library loads and exports every symbol include/abc_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "abc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from abc_b200 import _capi
    lib = _capi.load()
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libabc_b200.so lacks %s" % n
    # the Python binding covers the same set
    assert sorted(_capi.SYMBOLS) == names


def test_no_cpu_fallback_without_gpu():
    """Without a device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from abc_b200 import AbcError, CudaCiphertextFactory
    with pytest.raises(AbcError, match="no CUDA device|CUDA"):
        CudaCiphertextFactory(4096)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "abc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "bfv_oracle" not in txt and "liboracle" not in txt, f
