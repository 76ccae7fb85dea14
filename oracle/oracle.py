"""ctypes loader for the CPU oracle (oracle/s2k_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libs2k_oracle.so"

REGULAR, HPC, SIMD, HPCSIMD = 0, 1, 2, 3
NT1_32, NT2_31 = 0, 1

_lib = None


def build(force: bool = False) -> Path:
    srcs = [HERE / "s2k_oracle.c", HERE / "s2k_cpu_avx512.c", HERE / "Makefile"]
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.s2k_oracle_bound_scalar.restype = C.c_uint32
        L.s2k_oracle_bound_scalar.argtypes = [C.c_double]
        L.s2k_oracle_bound_simd.restype = C.c_uint32
        L.s2k_oracle_bound_simd.argtypes = [C.c_uint32]
        L.s2k_oracle_hpc.restype = C.c_size_t
        L.s2k_oracle_hpc.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.s2k_oracle_encode_rle.restype = C.c_size_t
        L.s2k_oracle_encode_rle.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.s2k_oracle_encode_rle_simd.restype = C.c_size_t
        L.s2k_oracle_encode_rle_simd.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.s2k_oracle_minimizers.restype = C.c_long
        L.s2k_oracle_minimizers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s2k_oracle_windows.restype = C.c_long
        L.s2k_oracle_windows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s2k_oracle_kminmers.restype = C.c_long
        L.s2k_oracle_kminmers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_long)]
        L.s2k_oracle_closed_minimizers.restype = C.c_long
        L.s2k_oracle_closed_minimizers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_uint64, C.c_int, C.c_long, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s2k_oracle_closed_windows.restype = C.c_long
        L.s2k_oracle_closed_windows.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_size_t]
        L.s2k_oracle_synth_word.restype = C.c_uint64
        L.s2k_oracle_synth_word.argtypes = [C.c_uint64, C.c_uint64]
        L.s2k_oracle_synth.restype = None
        L.s2k_oracle_synth.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.s2k_oracle_batch.restype = C.c_long
        L.s2k_oracle_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.s2k_cpu_has_avx512.restype = C.c_int
        L.s2k_cpu_has_avx512.argtypes = []
        L.s2k_cpu_avx512_batch.restype = C.c_long
        L.s2k_cpu_avx512_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p]
        L.s2k_oracle_digest_items.restype = None
        L.s2k_oracle_digest_items.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _seq(seq) -> np.ndarray:
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    if isinstance(seq, str):
        return np.frombuffer(seq.encode(), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def bound_scalar(density: float) -> int:
    return int(lib().s2k_oracle_bound_scalar(density))


def bound_simd(bound: int) -> int:
    return int(lib().s2k_oracle_bound_simd(bound))


def hpc(seq) -> bytes:
    s = _seq(seq)
    out = np.empty(max(len(s), 1), dtype=np.uint8)
    m = lib().s2k_oracle_hpc(_p(s), len(s), _p(out))
    return out[:m].tobytes()


def encode_rle(seq):
    s = _seq(seq)
    out = np.empty(max(len(s), 1), dtype=np.uint8)
    pos = np.empty(max(len(s), 1), dtype=np.uint64)
    m = lib().s2k_oracle_encode_rle(_p(s), len(s), _p(out), _p(pos))
    return out[:m].tobytes(), pos[:m].copy()


def encode_rle_simd(seq):
    s = _seq(seq)
    out = np.empty(max(len(s), 1), dtype=np.uint8)
    pos = np.empty(max(len(s), 1), dtype=np.uint32)
    m = lib().s2k_oracle_encode_rle_simd(_p(s), len(s), _p(out), _p(pos))
    return out[:m].tobytes(), pos[:m].copy()


def minimizers(seq, l: int, density: float, mode: int, variant: int = NT1_32):
    """-> (start u64[], end u64[], hash u32[]) for one sequence."""
    s = _seq(seq)
    cap = max(len(s), 1)
    st = np.empty(cap, dtype=np.uint64)
    en = np.empty(cap, dtype=np.uint64)
    h = np.empty(cap, dtype=np.uint32)
    n = lib().s2k_oracle_minimizers(_p(s), len(s), l, density, mode, variant, _p(st), _p(en), _p(h), cap)
    if n < 0:
        raise ValueError(f"oracle rejected parameters (code {n})")
    return st[:n].copy(), en[:n].copy(), h[:n].copy()


def kminmers(seq, l: int, k: int, density: float, mode: int, variant: int = NT1_32):
    """-> dict(hash u64[], start u64[], end u64[], offset u64[], rev u8[], n_minimizers) for one sequence."""
    s = _seq(seq)
    cap = max(len(s), 1)
    h = np.empty(cap, dtype=np.uint64)
    st = np.empty(cap, dtype=np.uint64)
    en = np.empty(cap, dtype=np.uint64)
    off = np.empty(cap, dtype=np.uint64)
    rv = np.empty(cap, dtype=np.uint8)
    nm = C.c_long(0)
    n = lib().s2k_oracle_kminmers(_p(s), len(s), l, k, density, mode, variant,
                                  _p(h), _p(st), _p(en), _p(off), _p(rv), cap, C.byref(nm))
    if n < 0:
        raise ValueError(f"oracle rejected parameters (code {n})")
    return dict(hash=h[:n].copy(), start=st[:n].copy(), end=en[:n].copy(), offset=off[:n].copy(),
                rev=rv[:n].copy(), n_minimizers=int(nm.value))


def closed_minimizers(seq, l, hpc_on, simd_tables, width, bound, strict, last_rule, end_rule):
    s = _seq(seq)
    cap = max(len(s), 1)
    st = np.empty(cap, dtype=np.uint64)
    en = np.empty(cap, dtype=np.uint64)
    h = np.empty(cap, dtype=np.uint64)
    n = lib().s2k_oracle_closed_minimizers(_p(s), len(s), l, int(hpc_on), int(simd_tables), width, bound,
                                           int(strict), last_rule, end_rule, _p(st), _p(en), _p(h), cap)
    return st[:n].copy(), en[:n].copy(), h[:n].copy()


def closed_windows(mhash, k, mix_u32=True):
    m = np.ascontiguousarray(mhash, dtype=np.uint64)
    cap = max(len(m), 1)
    h = np.empty(cap, dtype=np.uint64)
    rv = np.empty(cap, dtype=np.uint8)
    n = lib().s2k_oracle_closed_windows(_p(m), len(m), k, int(mix_u32), _p(h), _p(rv), cap)
    return h[:n].copy(), rv[:n].copy()


def closed_profile(seq, l, density, mode, variant=NT1_32):
    """Appendix-A.6 closed form for a HashMode (independent of the procedural restatement)."""
    b = bound_scalar(density)
    hpc_on = mode in (HPC, HPCSIMD)
    simd = mode in (SIMD, HPCSIMD)
    if simd:
        bs = bound_simd(b)
        if variant == NT2_31:
            return closed_minimizers(seq, l, hpc_on, True, 31, bs // 2, True, 0, 0)
        return closed_minimizers(seq, l, hpc_on, True, 32, bs, True, 2, 0)
    if mode == HPC:
        return closed_minimizers(seq, l, True, False, 32, b, False, 1, 1)
    return closed_minimizers(seq, l, False, False, 32, b, False, 0, 0)


def synth(seed: int, first: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint8)
    lib().s2k_oracle_synth(seed, first, count, _p(out))
    return out


def synth_word(seed: int, j: int) -> int:
    return int(lib().s2k_oracle_synth_word(seed, j))


def batch(bases, seq_off, l, k, density, mode, variant=NT1_32, threads=1, want_counts=True, want_digest=False):
    """Runs the oracle over a batch of reads on `threads` host threads (mirrors src/main.rs:65-79)."""
    b = _seq(bases)
    so = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = len(so) - 1
    km = np.zeros(n, dtype=np.uint64) if want_counts else None
    mc = np.zeros(n, dtype=np.uint64) if want_counts else None
    dg = np.zeros(n, dtype=np.uint64) if want_digest else None
    tm = C.c_uint64(0)
    tot = lib().s2k_oracle_batch(_p(b), _p(so), n, l, k, density, mode, variant, threads, _p(km), _p(mc), _p(dg),
                                 C.byref(tm))
    return dict(total=int(tot), total_min=int(tm.value), km_cnt=km, min_cnt=mc, digest=dg)


def has_avx512() -> bool:
    return bool(lib().s2k_cpu_has_avx512())


def avx512_batch(bases, seq_off, l, k, density, mode, threads=1, want_counts=True, want_digest=False):
    """AVX-512 restatement of the reference's Simd/HpcSimd path (oracle/s2k_cpu_avx512.c): the CPU baseline."""
    b = _seq(bases)                     # each read is copied into a padded per-thread buffer inside the C code
    so = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = len(so) - 1
    km = np.zeros(n, dtype=np.uint64) if want_counts else None
    dg = np.zeros(n, dtype=np.uint64) if want_digest else None
    tot = lib().s2k_cpu_avx512_batch(_p(b), _p(so), n, l, k, density, mode, threads, _p(km), _p(dg))
    if tot < 0:
        raise RuntimeError("AVX-512 baseline unavailable (CPU flags or mode)")
    return dict(total=int(tot), km_cnt=km, digest=dg)


def digest_items(hash_, start, end, rev, km_off) -> np.ndarray:
    km_off = np.ascontiguousarray(km_off, dtype=np.uint64)
    n = len(km_off) - 1
    out = np.zeros(n, dtype=np.uint64)
    lib().s2k_oracle_digest_items(_p(np.ascontiguousarray(hash_, dtype=np.uint64)),
                                  _p(np.ascontiguousarray(start, dtype=np.uint32)),
                                  _p(np.ascontiguousarray(end, dtype=np.uint32)),
                                  _p(np.ascontiguousarray(rev, dtype=np.uint8)), _p(km_off), n, _p(out))
    return out


def load_fixture(path=None) -> np.ndarray:
    """tests/golden/ecoli100k.2bit -> ASCII bases (see tests/golden/make_fixture.py)."""
    path = Path(path) if path else HERE.parent / "tests" / "golden" / "ecoli100k.2bit"
    raw = path.read_bytes()
    assert raw[:8] == b"S2KFIX01"
    (n,) = struct.unpack("<Q", raw[8:16])
    packed = np.frombuffer(raw[16:], dtype=np.uint8)
    idx = np.arange(n)
    codes = (packed[idx >> 2] >> (2 * (idx & 3)).astype(np.uint8)) & 3
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
