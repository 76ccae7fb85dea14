"""CPU tier: the C-ABI library loads and exports every symbol include/seq2kminmers.h declares; GPU-free entry
points behave; the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import re

import pytest


def test_library_exports_every_declared_symbol(S):
    header = S.HEADER_PATH.read_text()
    declared = set(re.findall(r"\b(s2k_[a-z_0-9]+)\s*\(", header))
    assert declared == set(S.ABI_SYMBOLS), declared ^ set(S.ABI_SYMBOLS)
    lib = ctypes.CDLL(str(S.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_gpu_free_entry_points(S, O):
    lib = S.Library()
    assert lib.c.s2k_abi_version() == 1
    assert lib.c.s2k_strerror(0) == b"ok" and lib.c.s2k_strerror(-2).startswith(b"l out of range")
    for d in (0.0, 1e-9, 0.0001, 0.001, 0.007, 0.01, 0.05, 0.1, 0.5, 0.999, 1.0, 2.0, -1.0):
        bs, bv, b31 = S.bounds(d, lib)
        assert bs == O.bound_scalar(d) and bv == O.bound_simd(bs) and b31 == bv // 2


def test_no_cpu_fallback_without_a_device(S):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.S2KError):
        S.Context(0)
    with pytest.raises(ImportError):
        S.Library("/nonexistent/libs2k_b200.so")
