import os
import importlib
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


os.environ.setdefault("S2K_WATCHDOG", "1")      # the library reports where a pipelined run stands every 5 s once it takes that long


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def S():
    """The product package (hyphenated directory)."""
    return importlib.import_module("rust-seq2kminmers_b200")


@pytest.fixture(scope="session")
def fixture_seq(O):
    return O.load_fixture()


@pytest.fixture(scope="session")
def emu_ctx(S):
    """Context over tests/emu/libs2k_emu.so: the product's CUDA sources compiled by g++ onto host threads.
    Test tier only -- checks kernel logic where there is no GPU; the product never loads it."""
    emu = ROOT / "tests" / "emu"
    subprocess.run([str(emu / "build_emu.sh")], check=True)
    ctx = S.Context(0, S.Library(emu / "libs2k_emu.so"))
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx(S):
    """Context over the nvcc-built product library on cuda:0.  Fails loudly if the library is missing."""
    ctx = S.Context(0)
    yield ctx
    ctx.close()


class Batches:
    """Seeded adversarial inputs shared by the emulation (CPU) and GPU parity tests."""

    def __init__(self, seed=12345):
        self.rng = np.random.default_rng(seed)

    def seq(self, n, alphabet=b"ACGT", runp=0.0):
        a = np.frombuffer(alphabet, dtype=np.uint8)
        s = a[self.rng.integers(0, len(a), int(n))]
        if runp > 0 and n > 0:
            s = np.repeat(s, self.rng.geometric(1 - runp, int(n)))[:int(n)]
        return s

    @staticmethod
    def pack(seqs):
        so = np.zeros(len(seqs) + 1, dtype=np.uint64)
        so[1:] = np.cumsum([len(s) for s in seqs])
        bases = np.concatenate(seqs) if int(so[-1]) > 0 else np.zeros(0, np.uint8)
        return bases, so

    def batch(self, lens, **kw):
        return self.pack([self.seq(n, **kw) for n in lens])


@pytest.fixture()
def batches():
    return Batches()


def assert_batch_matches_oracle(O, got, bases, so, l, k, d, mode, variant=0, check_minimizers=True):
    """Full-tuple, per-sequence comparison of a KminmersBatch with the oracle."""
    n = len(so) - 1
    for i in range(n):
        s = bases[int(so[i]):int(so[i + 1])]
        want = O.kminmers(s, l, k, d, int(mode), int(variant))
        a, b = int(got.km_off[i]), int(got.km_off[i + 1])
        ctxt = f"seq {i} len {len(s)} mode {int(mode)} variant {int(variant)} l {l} k {k} d {d}"
        assert b - a == len(want["hash"]), f"item count {b - a} != {len(want['hash'])}: {ctxt}"
        assert np.array_equal(got.hash[a:b], want["hash"]), "hash: " + ctxt
        assert np.array_equal(got.start[a:b], want["start"].astype(np.uint32)), "start: " + ctxt
        assert np.array_equal(got.end[a:b], want["end"].astype(np.uint32)), "end: " + ctxt
        assert np.array_equal(got.rev[a:b], want["rev"]), "rev: " + ctxt
        if check_minimizers and got.minimizers is not None:
            ms, me, mh = O.minimizers(s, l, d, int(mode), int(variant))
            g = got.minimizers_of(i)
            assert len(g) == len(mh), f"minimizer count {len(g)} != {len(mh)}: {ctxt}"
            assert np.array_equal(g["hash"], mh) and np.array_equal(g["start"], ms.astype(np.uint32)) \
                and np.array_equal(g["end"], me.astype(np.uint32)) and np.all(g["seq"] == i), "minimizers: " + ctxt
    assert int(got.km_off[0]) == 0 and int(got.km_off[-1]) == got.n_items
