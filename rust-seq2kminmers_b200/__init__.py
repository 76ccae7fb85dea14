"""rust-seq2kminmers_b200 -- B200-native sequence -> k-min-mer path, host-side mirror of the reference API.

The reference crate exposes (src/lib.rs:5-39, 70-132, 179-270; src/kminmer.rs:128-177):

    KminmersIterator::new(seq, l, k, density, mode) -> Iterator<Item = KminmerHash>
    HashMode::{Regular, Hpc, Simd, HpcSimd}
    KminmerHash { hash, start, end, offset, rev }            (equality/order on .hash only)
    NtHashHPCIterator / NtHashSIMDIterator / NtHashHPCSIMDIterator   -> (start, end, hash) minimizers
    hpc(), encode_rle(), encode_rle_simd()

This module keeps those names and argument meanings on top of the C ABI in include/seq2kminmers.h
(libs2k_b200.so: hand-written CUDA for sm_100a).  There is NO CPU implementation here: if the CUDA library
is missing or no GPU is present, constructing a Context raises.  (The directory name contains a hyphen, as
the reference crate's does; import it with importlib.import_module("rust-seq2kminmers_b200") or through the
`seq2kminmers_b200` alias module at the repository root.)
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from pathlib import Path
from typing import Iterator, Optional, Sequence

import numpy as np

__all__ = [
    "HashMode", "HashVariant", "KminmerHash", "KminmerVec", "Kminmer", "KminmersIterator", "KminmersBatch", "Context", "Library",
    "S2KError", "NtHashHPCIterator", "NtHashSIMDIterator", "NtHashHPCSIMDIterator", "hpc", "encode_rle",
    "encode_rle_simd", "bounds", "default_library", "LIB_PATH",
]

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libs2k_b200.so"
HEADER_PATH = PKG_DIR.parent / "include" / "seq2kminmers.h"


class HashMode(enum.IntEnum):
    """src/lib.rs:21-27."""
    Regular = 0
    Hpc = 1
    Simd = 2
    HpcSimd = 3


class HashVariant(enum.IntEnum):
    """NT1_32: src/nthash_avx512_32.rs / src/nthash_hpc.rs; NT2_31: src/nthash2_avx512_32.rs; NT1_64: the crate built
    with `pub type H = u64` (src/lib.rs:30-32; modes Regular and Hpc; golden vector tests/main.rs:18-39); NT1_16: the
    crate built with `pub type H = u16` (src/lib.rs:29; modes Regular and Hpc; parity unpinned)."""
    NT1_32 = 0
    NT2_31 = 1
    NT1_64 = 2
    NT1_16 = 3


class S2KError(RuntimeError):
    """Raised where the reference panics (unwrap/assert) or on a CUDA failure; carries the C status code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"s2k status {status}: {message}")
        self.status = status


class _Params(C.Structure):
    _fields_ = [("l", C.c_uint32), ("k", C.c_uint32), ("density", C.c_double), ("mode", C.c_int32),
                ("variant", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("n_items", C.c_uint64), ("n_minimizers", C.c_uint64),
                ("hash", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p), ("rev", C.c_void_p),
                ("km_off", C.c_void_p), ("minimizers", C.c_void_p), ("min_off", C.c_void_p), ("min_cnt", C.c_void_p),
                ("location", C.c_int32), ("reserved", C.c_int32)]


class _RleResult(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("n_hpc", C.c_uint64), ("hpc", C.c_void_p), ("pos", C.c_void_p),
                ("hpc_off", C.c_void_p), ("location", C.c_int32), ("reserved", C.c_int32)]


class _CountResult(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("n_items", C.c_uint64), ("hash", C.c_void_p), ("count", C.c_void_p),
                ("first", C.c_void_p), ("location", C.c_int32), ("reserved", C.c_int32)]


MINIMIZER_DTYPE = np.dtype([("hash", "<u4"), ("start", "<u4"), ("end", "<u4"), ("seq", "<u4")])

# every symbol include/seq2kminmers.h declares
ABI_SYMBOLS = (
    "s2k_ctx_create", "s2k_ctx_destroy", "s2k_ctx_set_flags", "s2k_run", "s2k_run_device", "s2k_encode_rle",
    "s2k_bounds", "s2k_host_alloc", "s2k_host_free", "s2k_last_error", "s2k_strerror", "s2k_abi_version",
    "s2k_launch_count", "s2k_ctx_set_timing", "s2k_last_kernel_ms", "s2k_synth_device",
    "s2k_ctx_set_slab_bytes", "s2k_run_fastx", "s2k_last_fastx", "s2k_ctx_set_transport",
    "s2k_last_transport", "s2k_run_packed2", "s2k_pack2", "s2k_count_device", "s2k_count_partition_device", "s2k_count_part",
    "s2k_bound_u64", "s2k_last_minimizer_hash_hi", "s2k_bound_u16",
)


class Library:
    """The C ABI, loaded from a shared object.  The default is the nvcc-built product library."""

    def __init__(self, path: Optional[Path] = None):
        self.path = Path(path) if path else LIB_PATH
        if not self.path.exists():
            raise ImportError(
                f"{self.path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for this path.")
        self.c = L = C.CDLL(str(self.path))
        vp = C.c_void_p
        L.s2k_ctx_create.restype = C.c_int
        L.s2k_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.s2k_ctx_destroy.restype = None
        L.s2k_ctx_destroy.argtypes = [vp]
        L.s2k_ctx_set_flags.restype = C.c_int
        L.s2k_ctx_set_flags.argtypes = [vp, C.c_uint32]
        L.s2k_run.restype = C.c_int
        L.s2k_run.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(_Params), C.POINTER(_Result)]
        L.s2k_run_device.restype = C.c_int
        L.s2k_run_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(_Params), vp, C.POINTER(_Result)]
        L.s2k_encode_rle.restype = C.c_int
        L.s2k_encode_rle.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(_RleResult)]
        L.s2k_bound_u64.restype = C.c_uint64
        L.s2k_bound_u64.argtypes = [C.c_double]
        L.s2k_bound_u16.restype = C.c_uint32
        L.s2k_bound_u16.argtypes = [C.c_double]
        L.s2k_last_minimizer_hash_hi.restype = C.c_int
        L.s2k_last_minimizer_hash_hi.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.s2k_bounds.restype = None
        L.s2k_bounds.argtypes = [C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.s2k_host_alloc.restype = C.c_int
        L.s2k_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.s2k_host_free.restype = None
        L.s2k_host_free.argtypes = [vp]
        L.s2k_last_error.restype = C.c_char_p
        L.s2k_last_error.argtypes = [vp]
        L.s2k_strerror.restype = C.c_char_p
        L.s2k_strerror.argtypes = [C.c_int]
        L.s2k_abi_version.restype = C.c_int
        L.s2k_abi_version.argtypes = []
        L.s2k_launch_count.restype = C.c_uint64
        L.s2k_launch_count.argtypes = [vp]
        L.s2k_ctx_set_timing.restype = C.c_int
        L.s2k_ctx_set_timing.argtypes = [vp, C.c_int]
        L.s2k_ctx_set_slab_bytes.restype = C.c_int
        L.s2k_ctx_set_slab_bytes.argtypes = [vp, C.c_uint64]
        L.s2k_run_packed2.restype = C.c_int
        L.s2k_run_packed2.argtypes = L.s2k_run.argtypes
        L.s2k_pack2.restype = C.c_int64
        L.s2k_pack2.argtypes = [vp, C.c_uint64, vp, C.c_int]
        L.s2k_last_transport.restype = C.c_int
        L.s2k_last_transport.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.s2k_ctx_set_transport.restype = C.c_int
        L.s2k_ctx_set_transport.argtypes = [vp, C.c_int, C.c_double]
        L.s2k_run_fastx.restype = C.c_int
        L.s2k_run_fastx.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(_Params), C.POINTER(_Result)]
        L.s2k_last_fastx.restype = C.c_int
        L.s2k_last_fastx.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp)]
        L.s2k_synth_device.restype = C.c_int
        L.s2k_synth_device.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]
        L.s2k_last_kernel_ms.restype = C.c_int
        L.s2k_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
        L.s2k_count_device.restype = C.c_int
        L.s2k_count_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, vp, C.POINTER(_CountResult)]
        L.s2k_count_partition_device.restype = C.c_int
        L.s2k_count_partition_device.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, vp, vp]
        L.s2k_count_part.restype = C.c_uint32
        L.s2k_count_part.argtypes = [C.c_uint64, C.c_uint32]


_default: Optional[Library] = None


def default_library() -> Library:
    global _default
    if _default is None:
        _default = Library()
    return _default


def bounds(density: float, lib: Optional[Library] = None):
    """(bound_scalar, bound_simd, bound_31) -- src/lib.rs:91, src/nthash_avx512_32.rs:47-48, nthash2:52-54."""
    lib = lib or default_library()
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.c.s2k_bounds(float(density), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.empty(0, dtype=dtype)
    dt = np.dtype(dtype)
    buf = (C.c_uint8 * (count * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=count)


@dataclass
class KminmersBatch:
    """SoA result of one batched run (host copies).  Items of sequence r: [km_off[r], km_off[r+1])."""
    n_seqs: int
    hash: np.ndarray      # u64 [n_items]
    start: np.ndarray     # u32
    end: np.ndarray       # u32
    rev: np.ndarray       # u8
    km_off: np.ndarray    # u64 [n_seqs+1]
    min_off: np.ndarray   # u64 [n_seqs+1]
    min_cnt: np.ndarray   # u32 [n_seqs]
    n_minimizers: int
    minimizers: Optional[np.ndarray] = None   # MINIMIZER_DTYPE [n_minimizers]

    @property
    def n_items(self) -> int:
        return int(self.hash.shape[0])

    def items(self, r: int) -> Iterator["KminmerHash"]:
        a, b = int(self.km_off[r]), int(self.km_off[r + 1])
        for i in range(a, b):
            yield KminmerHash(int(self.hash[i]), int(self.start[i]), int(self.end[i]), i - a, bool(self.rev[i]))

    def kminmer_vecs(self, r: int, k: int) -> Iterator["KminmerVec"]:
        """The items of sequence r in the `KminmerVec` flavour (src/kminmer.rs:17-38): every window of k consecutive
        minimizers as the vector of their hashes, start of the first, end of the last, offset.  Needs the minimizer
        stream (want_minimizers=True)."""
        m = self.minimizers_of(r)
        for c in range(max(0, len(m) - k + 1)):
            w = m[c:c + k]
            yield KminmerVec(w["hash"], int(w["start"][0]), int(w["end"][-1]), c)

    def minimizers_of(self, r: int) -> np.ndarray:
        """Minimizers of sequence r that feed the window stage (what the reference's inner iterator yields)."""
        if self.minimizers is None:
            raise ValueError("run with want_minimizers=True")
        a = int(self.min_off[r])
        return self.minimizers[a:a + int(self.min_cnt[r])]


@dataclass(frozen=True)
class KminmerHash:
    """src/kminmer.rs:128-135; equality and order by `hash` only (src/kminmer.rs:181-203)."""
    hash: int
    start: int
    end: int
    offset: int
    rev: bool

    def get_hash(self) -> int:   # trait Kminmer, src/kminmer.rs:12-15
        return self.hash

    def __eq__(self, other):
        return isinstance(other, KminmerHash) and self.hash == other.hash

    def __lt__(self, other):
        return self.hash < other.hash

    def __hash__(self):
        return hash(self.hash)


Kminmer = KminmerHash  # KminmerType, src/lib.rs:39


def fxhash64_u32_slice(mers) -> int:
    """`fxhash::hash64(&Vec<u32>)` (crate fxhash 0.2.1, `Cargo.toml` of the reference; not under /root/reference):
    FxHasher64 over the slice's `Hash` impl = `write_usize(len)`, then the elements' bytes in one `write` -- eight bytes
    per step, then four: state = (rotl(state, 5) ^ word) * 0x517cc1b727220a95.  PARITY UNPINNED: the reference holds no
    vector for it (only `KminmerHash::new`, which the iterator does not call, uses it: src/kminmer.rs:138-161)."""
    m64, seed = (1 << 64) - 1, 0x517cc1b727220a95
    h = 0

    def word(h, w):
        return ((((h << 5) | (h >> 59)) & m64) ^ w) * seed & m64
    v = [int(x) & 0xffffffff for x in mers]
    h = word(h, len(v))
    for i in range(0, len(v) - 1, 2):
        h = word(h, v[i] | (v[i + 1] << 32))
    if len(v) & 1:
        h = word(h, v[-1])
    return h


class KminmerVec:
    """src/kminmer.rs:17-126: a k-min-mer as the vector of its minimizer hashes, stored in canonical orientation
    (`normalize`: the reversed vector if it is lexicographically smaller, then rev = True); equality and order by the
    vector.  Built on the host from the minimizer stream of a batch (KminmersBatch.kminmer_vecs)."""

    def __init__(self, mers, start: int, end: int, offset: int):       # Kminmer::new, src/kminmer.rs:27-38
        self._mers = [int(x) for x in mers]
        self.start, self.end, self.offset, self.rev = int(start), int(end), int(offset), False
        self.normalize()

    def normalize(self):                                               # src/kminmer.rs:53-60
        r = self._mers[::-1]
        if r < self._mers:
            self._mers, self.rev = r, True

    def is_normalized(self) -> bool:                                   # src/kminmer.rs:63-67
        return self._mers <= self._mers[::-1]

    def mers(self):                                                    # src/kminmer.rs:80-82
        return list(self._mers)

    def print(self) -> str:                                            # src/kminmer.rs:70-77: first two digits of each hash
        return "".join(str(x)[:2] + " " for x in self._mers)

    def get_hash_u64(self) -> int:                                     # src/kminmer.rs:90-92
        return fxhash64_u32_slice(self._mers)

    def to_kminmer_hash(self) -> "KminmerHash":                        # KminmerHash::new, src/kminmer.rs:138-161
        return KminmerHash(self.get_hash_u64(), self.start, self.end, self.offset, self.rev)

    def __eq__(self, other):
        return isinstance(other, KminmerVec) and self._mers == other._mers

    def __lt__(self, other):
        return self._mers < other._mers

    def __hash__(self):
        return hash(tuple(self._mers))

    def __repr__(self):
        return f"KminmerVec(mers={self._mers}, start={self.start}, end={self.end}, offset={self.offset}, rev={self.rev})"


def _as_u8(seq) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


class Context:
    """One s2k_ctx: one CUDA device, one stream, grow-only buffers.  Single-threaded, like one iterator."""

    def __init__(self, device: int = 0, lib: Optional[Library] = None):
        self.lib = lib or default_library()
        h = C.c_void_p()
        st = self.lib.c.s2k_ctx_create(int(device), C.byref(h))
        if st != 0:
            raise S2KError(st, self.lib.c.s2k_strerror(st).decode() +
                           " (s2k_ctx_create: is a CUDA device visible? this path has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.c.s2k_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st: int):
        if st != 0:
            msg = self.lib.c.s2k_last_error(self.h).decode() or self.lib.c.s2k_strerror(st).decode()
            raise S2KError(st, msg)

    # -- host buffers in, host (pinned) results out ------------------------------------------------------
    def run(self, bases, seq_off, l: int, k: int, density: float, mode: HashMode,
            variant: HashVariant = HashVariant.NT1_32, want_minimizers: bool = False, copy: bool = True,
            no_tail_rule: bool = False, debug_tiny_cap: bool = False,
            packed2: bool = False) -> KminmersBatch:
        """packed2=True: `bases` is 2-bit packed input (see pack2 / s2k_run_packed2), seq_off still counts bases."""
        b = _as_u8(bases)
        so = np.ascontiguousarray(seq_off, dtype=np.uint64)
        if so.ndim != 1 or so.shape[0] < 1:
            raise ValueError("seq_off must hold n_seqs+1 offsets")
        n = so.shape[0] - 1
        if (int(so[-1]) + 3) // 4 > b.shape[0] if packed2 else int(so[-1]) > b.shape[0]:
            raise ValueError("seq_off[-1] exceeds len(bases)")
        self._check(self.lib.c.s2k_ctx_set_flags(self.h, (1 if want_minimizers else 0) | (2 if no_tail_rule else 0) |
                                                 (8 if debug_tiny_cap else 0)))
        p = _Params(int(l), int(k), float(density), int(mode), int(variant))
        r = _Result()
        try:
            fn = self.lib.c.s2k_run_packed2 if packed2 else self.lib.c.s2k_run
            self._check(fn(self.h, b.ctypes.data, so.ctypes.data, n, C.byref(p), C.byref(r)))
        finally:
            self.lib.c.s2k_ctx_set_flags(self.h, 0)
        f = (lambda a: a.copy()) if copy else (lambda a: a)
        mins = f(_view(r.minimizers, r.n_minimizers, MINIMIZER_DTYPE)) if want_minimizers else None
        return KminmersBatch(n, f(_view(r.hash, r.n_items, np.uint64)), f(_view(r.start, r.n_items, np.uint32)),
                             f(_view(r.end, r.n_items, np.uint32)), f(_view(r.rev, r.n_items, np.uint8)),
                             f(_view(r.km_off, n + 1, np.uint64)), f(_view(r.min_off, n + 1, np.uint64)),
                             f(_view(r.min_cnt, n, np.uint32)), int(r.n_minimizers), mins)

    def pack2(self, bases, host_threads: int = 8):
        """ASCII A/C/G/T -> 2-bit packed bytes for run(..., packed2=True) (s2k_pack2).  Raises on any other byte."""
        b = _as_u8(bases)
        out = np.zeros((b.shape[0] + 3) // 4 + 8, dtype=np.uint8)
        st = self.lib.c.s2k_pack2(b.ctypes.data, b.shape[0], out.ctypes.data, int(host_threads))
        if st != 0:
            raise ValueError("pack2: input holds bytes other than upper-case A/C/G/T")
        return out

    # -- device buffers in, device results out ------------------------------------------------------------
    def run_device(self, d_bases_ptr: int, d_seq_off_ptr: int, n_seqs: int, n_bases: int, l: int, k: int,
                   density: float, mode: HashMode, variant: HashVariant = HashVariant.NT1_32, stream: int = 0,
                   no_tail_rule: bool = False, no_minimizer_stream: bool = False) -> _Result:
        """Raw device-pointer form (pointers as ints).  Returns the ctypes result struct (device pointers).
        no_minimizer_stream (S2K_NO_MINIMIZER_STREAM): result.minimizers is NULL, the window stage reads the minimizer
        records in place (one pass over them less); everything else in the result is unchanged."""
        p = _Params(int(l), int(k), float(density), int(mode), int(variant))
        r = _Result()
        self._check(self.lib.c.s2k_ctx_set_flags(self.h, (2 if no_tail_rule else 0) | (16 if no_minimizer_stream else 0)))
        try:
            self._check(self.lib.c.s2k_run_device(self.h, C.c_void_p(d_bases_ptr), C.c_void_p(d_seq_off_ptr), int(n_seqs),
                                                  int(n_bases), C.byref(p), C.c_void_p(stream), C.byref(r)))
        finally:
            self.lib.c.s2k_ctx_set_flags(self.h, 0)
        return r

    def run_fastx(self, path, nb_threads: int, l: int, k: int, density: float, mode: HashMode,
                  variant: HashVariant = HashVariant.NT1_32, copy: bool = True, keep_bases: bool = True):
        """parallel_fastx(path, nb_threads, |seq, id| KminmersIterator::new(seq, l, k, density, mode)) of src/main.rs:65-79.
        Returns (KminmersBatch, bases u8[], seq_off u64[]) -- the parsed records in file order and their k-min-mers.
        keep_bases=False: large files stream through the device slab by slab (packed on the fly, never materialised on the
        host); `bases` is then None."""
        p = _Params(int(l), int(k), float(density), int(mode), int(variant))
        r = _Result()
        self._check(self.lib.c.s2k_ctx_set_flags(self.h, 32 if keep_bases else 0))     # S2K_FASTX_KEEP_BASES
        try:
            self._check(self.lib.c.s2k_run_fastx(self.h, str(path).encode(), int(nb_threads), C.byref(p), C.byref(r)))
        finally:
            self.lib.c.s2k_ctx_set_flags(self.h, 0)
        ns, nb, pb, po = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.c.s2k_last_fastx(self.h, C.byref(ns), C.byref(nb), C.byref(pb), C.byref(po)))
        f = (lambda a: a.copy()) if copy else (lambda a: a)
        n = int(ns.value)
        batch = KminmersBatch(n, f(_view(r.hash, r.n_items, np.uint64)), f(_view(r.start, r.n_items, np.uint32)),
                              f(_view(r.end, r.n_items, np.uint32)), f(_view(r.rev, r.n_items, np.uint8)),
                              f(_view(r.km_off, n + 1, np.uint64)), f(_view(r.min_off, n + 1, np.uint64)),
                              f(_view(r.min_cnt, n, np.uint32)), int(r.n_minimizers), None)
        bases = f(_view(pb.value, int(nb.value), np.uint8)) if pb.value else None      # None: the file was streamed
        return batch, bases, f(_view(po.value, n + 1, np.uint64))

    def count_device(self, d_hash_ptr: int, n_items: int, d_id_ptr: int = 0, id_base: int = 0, stream: int = 0) -> _CountResult:
        """s2k_count_device: abundance of the distinct k-min-mer hashes of a device-resident item stream (the consumer side,
        rust-mdbg's map keyed by KminmerHash).  Returns the ctypes result (device pointers: hash u64, count u32, first u64)."""
        r = _CountResult()
        self._check(self.lib.c.s2k_count_device(self.h, C.c_void_p(d_hash_ptr), C.c_void_p(d_id_ptr or None), int(n_items),
                                                int(id_base), C.c_void_p(stream), C.byref(r)))
        return r

    def count_partition_device(self, d_hash_ptr: int, n_items: int, id_base: int, n_parts: int, d_out_hash_ptr: int,
                               d_out_id_ptr: int, stream: int = 0) -> np.ndarray:
        """s2k_count_partition_device: items bucketed by the rank that counts their hash; returns the bucket sizes."""
        counts = np.zeros(int(n_parts), dtype=np.uint64)
        self._check(self.lib.c.s2k_count_partition_device(self.h, C.c_void_p(d_hash_ptr), int(n_items), int(id_base), int(n_parts),
                                                          counts.ctypes.data, C.c_void_p(d_out_hash_ptr), C.c_void_p(d_out_id_ptr),
                                                          C.c_void_p(stream)))
        return counts

    def synth_device(self, seed: int, first: int, count: int, d_out_ptr: int, stream: int = 0):
        """Fill device memory with the synthetic base stream of SURVEY.md 8(d)."""
        self._check(self.lib.c.s2k_synth_device(self.h, int(seed), int(first), int(count), C.c_void_p(d_out_ptr),
                                                C.c_void_p(stream)))

    def encode_rle(self, bases, seq_off, scalar_rule: bool = False):
        """Batched encode_rle_simd (src/hpc.rs:44); scalar_rule=True (S2K_RLE_SCALAR_RULE): the rule of the scalar
        encode_rle (src/hpc.rs:14), where only repeated bytes out of "ACTGactgNn" are dropped."""
        b = _as_u8(bases)
        so = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = so.shape[0] - 1
        r = _RleResult()
        self._check(self.lib.c.s2k_ctx_set_flags(self.h, 4 if scalar_rule else 0))
        try:
            self._check(self.lib.c.s2k_encode_rle(self.h, b.ctypes.data, so.ctypes.data, n, C.byref(r)))
        finally:
            self.lib.c.s2k_ctx_set_flags(self.h, 0)
        return (_view(r.hpc, r.n_hpc, np.uint8).copy(), _view(r.pos, r.n_hpc, np.uint32).copy(),
                _view(r.hpc_off, n + 1, np.uint64).copy())

    def set_slab_bytes(self, nbytes: int):
        """Target slab size of the pipelined host path (0 = default 256 MiB)."""
        self._check(self.lib.c.s2k_ctx_set_slab_bytes(self.h, int(nbytes)))

    def set_transport(self, host_threads: int = 0, pack_ratio: float = 0.7):
        """2-bit transport of large host batches: share of slabs packed by host threads (0 = plain ASCII only)."""
        self._check(self.lib.c.s2k_ctx_set_transport(self.h, int(host_threads), float(pack_ratio)))

    def last_transport(self):
        """(bytes copied host->device, slabs packed, slabs plain) of the last host-buffer run."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.c.s2k_last_transport(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_timing(self, enabled: bool):
        self._check(self.lib.c.s2k_ctx_set_timing(self.h, 1 if enabled else 0))

    def last_kernel_ms(self):
        a, b, n = C.c_double(), C.c_double(), C.c_uint32()
        self._check(self.lib.c.s2k_last_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    @property
    def launch_count(self) -> int:
        return int(self.lib.c.s2k_launch_count(self.h))


class DeviceArray:
    """Zero-copy view of context-owned device memory (``__cuda_array_interface__``), e.g. for
    ``torch.as_tensor(DeviceArray(ptr, n, "<u8"), device="cuda")``.  Valid until the next run on the context."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr or 0), False),
                                         "version": 2, "strides": None}


_ctx: Optional[Context] = None


def _default_ctx() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


class KminmersIterator:
    """KminmersIterator::new(seq, l, k, density, mode) (src/lib.rs:89) as a Python iterator of KminmerHash.

    The whole sequence is processed on the GPU at construction (a batch of one); iteration then yields the
    items in the reference's order.  For throughput use Context.run on many sequences at once."""

    def __init__(self, seq, l: int, k: int, density: float, mode: HashMode,
                 variant: HashVariant = HashVariant.NT1_32, ctx: Optional[Context] = None):
        s = _as_u8(seq)
        ctx = ctx or _default_ctx()
        self.batch = ctx.run(s, np.array([0, s.shape[0]], dtype=np.uint64), l, k, density, mode, variant)
        self._it = self.batch.items(0)

    def __iter__(self):
        return self

    def __next__(self) -> KminmerHash:
        return next(self._it)


class _MinimizerIterator:
    mode = HashMode.Regular

    def __init__(self, seq, l: int, hash_bound_density: float, ctx: Optional[Context] = None,
                 variant: HashVariant = HashVariant.NT1_32):
        s = _as_u8(seq)
        ctx = ctx or _default_ctx()
        b = ctx.run(s, np.array([0, s.shape[0]], dtype=np.uint64), l, 1, hash_bound_density, self.mode, variant,
                    want_minimizers=True)
        self.items = b.minimizers_of(0)
        self._i = 0

    def __iter__(self):
        return self


class NtHashHPCIterator(_MinimizerIterator):
    """src/nthash_hpc.rs:193: Item = (start, end, hash).  Takes the density (the bound is derived as lib.rs:91)."""
    mode = HashMode.Hpc

    def __next__(self):
        if self._i >= len(self.items):
            raise StopIteration
        m = self.items[self._i]
        self._i += 1
        return int(m["start"]), int(m["end"]), int(m["hash"])


class NtHashHPCSIMDIterator(NtHashHPCIterator):
    """src/nthash_hpc_simd.rs:59: Item = (start, end, hash)."""
    mode = HashMode.HpcSimd


class NtHashSIMDIterator(_MinimizerIterator):
    """src/nthash_avx512_32.rs:82: Item = (pos, hash)."""
    mode = HashMode.Simd

    def __next__(self):
        if self._i >= len(self.items):
            raise StopIteration
        m = self.items[self._i]
        self._i += 1
        return int(m["start"]), int(m["hash"])


def encode_rle_simd(seq, ctx: Optional[Context] = None):
    """src/hpc.rs:44: (hpc bytes, run-start positions u32)."""
    s = _as_u8(seq)
    ctx = ctx or _default_ctx()
    h, p, _ = ctx.encode_rle(s, np.array([0, s.shape[0]], dtype=np.uint64))
    return h.tobytes(), p


def encode_rle(seq, ctx: Optional[Context] = None):
    """src/hpc.rs:7-25: runs collapse only for bytes out of "ACTGactgNn" (src/hpc.rs:14); positions widened to u64 like
    Vec<usize>."""
    s = _as_u8(seq)
    ctx = ctx or _default_ctx()
    h, p, _ = ctx.encode_rle(s, np.array([0, s.shape[0]], dtype=np.uint64), scalar_rule=True)
    return h.tobytes(), p.astype(np.uint64)


def hpc(seq, ctx: Optional[Context] = None) -> bytes:
    """src/hpc.rs:28."""
    return encode_rle_simd(seq, ctx)[0]


# ------------------------------------------------------------------------------------------------ minimizer TSV
# src/old/kminmers-readwrite.rs ("written by Baris but unused so far"): the minimizers of one sequence persisted as
# `position \t hash` lines in `{prefix}-{l}-{density}.mers`, and the k-min-mers (KminmerVec flavour) rebuilt from that
# file without the sequence.  Host side only; the minimizers themselves come from the device.
def _rust_display_f64(x: float) -> str:
    """`format!("{}", x)` for an f64: shortest round-trip digits, never an exponent, no trailing ".0"."""
    import decimal
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = format(decimal.Decimal(repr(x)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return "-0" if s in ("-0", "-") else s


def mers_path(prefix: str, l: int, density: float) -> str:
    """src/old/kminmers-readwrite.rs:21,120."""
    return f"{prefix}-{int(l)}-{_rust_display_f64(density)}.mers"


class KminmersWriteIterator:
    """src/old/kminmers-readwrite.rs:5-104: iterate the k-min-mers (KminmerVec) of `seq` and leave its minimizers in
    `{prefix}-{l}-{density}.mers`, one `position \\t hash` line each.  hpc=True runs mode Hpc (position = start of the
    minimizer in original coordinates), hpc=False mode Regular, whose positions the reference writes 1-based
    (`self.seq_pos += 1; j = self.seq_pos`, :88-89).  The device yields all minimizers of the sequence at once, so the
    file is complete when the constructor returns (the reference appends while it iterates); selection follows the
    chosen `variant` (the historical file used a u64 bound with `<`)."""

    def __init__(self, seq, l: int, k: int, density: float, hpc: bool, prefix: str, ctx: Optional["Context"] = None,
                 variant: HashVariant = HashVariant.NT1_32):
        s = _as_u8(seq)
        ctx = ctx or _default_ctx()
        self.l, self.k, self.path = int(l), int(k), mers_path(prefix, l, density)
        mode = HashMode.Hpc if hpc else HashMode.Regular
        b = ctx.run(s, np.array([0, s.shape[0]], dtype=np.uint64), l, k, density, mode, variant, want_minimizers=True)
        m = b.minimizers_of(0)
        self._pos = m["start"].astype(np.int64) + (0 if hpc else 1)
        self._hash = m["hash"].astype(np.uint64)
        with open(self.path, "w") as f:                                # truncate(true), :22-27
            f.write("".join(f"{int(p)}\t{int(h)}\n" for p, h in zip(self._pos, self._hash)))
        self._c = 0

    def __iter__(self):
        return self

    def __next__(self) -> "KminmerVec":
        c, k = self._c, self.k
        if c + k > len(self._pos):
            raise StopIteration
        self._c += 1
        return KminmerVec(self._hash[c:c + k], int(self._pos[c]), int(self._pos[c + k - 1]) + self.l - 1, c)   # :93


class KminmersReadIterator:
    """src/old/kminmers-readwrite.rs:107-165: the same k-min-mers from the `.mers` file alone."""

    def __init__(self, l: int, k: int, density: float, prefix: str):
        self.l, self.k = int(l), int(k)
        self._f = open(mers_path(prefix, l, density), "r")             # "Could not open minimizer index."
        self._pos, self._sk, self._count = [], [], 0

    def __iter__(self):
        return self

    def __next__(self) -> "KminmerVec":
        while True:
            line = self._f.readline()
            if not line:
                self._f.close()
                raise StopIteration
            v = line.rstrip("\n").split("\t")
            self._pos.append(int(v[0]))
            self._sk.append(int(v[1]))
            if len(self._sk) == self.k:
                km = KminmerVec(self._sk, self._pos[0], self._pos[self.k - 1] + self.l - 1, self._count)
                self._sk, self._pos = self._sk[1:], self._pos[1:]
                self._count += 1
                return km
