"""CPU tier: the product's CUDA sources, host-emulated (tests/emu) and built with AddressSanitizer, on an adversarial
batch with exact-size input buffers.  compute-sanitizer is closed on the GPU pool, so this is where out-of-bounds
accesses in kernel logic (shared-memory arrays live on the heap in the emulation) and over-reads of the caller's
buffers would show up."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_emulated_kernels_are_asan_clean(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not Path(asan).exists():
        pytest.skip("libasan not available")
    lib = tmp_path / "libs2k_emu_asan.so"
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-fPIC", "-shared",
                    "-pthread", "-DS2K_EMU", f"-I{ROOT / 'tests' / 'emu'}", "-x", "c++",
                    str(ROOT / "rust-seq2kminmers_b200" / "csrc" / "s2k_api.cu"), "-o", str(lib)], check=True)
    env = dict(os.environ, S2K_ROOT=str(ROOT), S2K_ASAN_LIB=str(lib), LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "emu" / "asan_workload.py")], capture_output=True, text=True,
                         env=env, timeout=1500)
    assert out.returncode == 0 and "asan workload done" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    assert "AddressSanitizer" not in out.stderr
