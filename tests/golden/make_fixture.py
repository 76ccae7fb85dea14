"""Regenerates tests/golden/ecoli100k.2bit from the reference's own test fixture.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_fixture.py

The reference's integration test (tests/main.rs:15-16) uses line index 1 of
tests/ecoli.genome.100k.fa (99 925 upper-case ACGT bases, single record).  We keep that
sequence as test DATA in a 2-bit packed form: 16-byte header ("S2KFIX01", u64 LE length),
then 4 bases per byte, base i in bits 2*(i%4).. of byte i//4, A=0 C=1 G=2 T=3.
"""
import struct
import sys
from pathlib import Path

SRC = Path("/root/reference/tests/ecoli.genome.100k.fa")
DST = Path(__file__).with_name("ecoli100k.2bit")


def main() -> int:
    seq = SRC.read_text().split("\n")[1].encode()
    assert set(seq) <= set(b"ACGT"), "fixture is expected to be pure upper-case ACGT"
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    out = bytearray((len(seq) + 3) // 4)
    for i, b in enumerate(seq):
        out[i >> 2] |= code[b] << (2 * (i & 3))
    DST.write_bytes(b"S2KFIX01" + struct.pack("<Q", len(seq)) + bytes(out))
    print(f"wrote {DST} ({len(seq)} bases)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
