"""Twin of the reference's driver (src/main.rs): `python -m seq2kminmers_b200 <fasta|fastq> <nb_threads> [mode]`.
No argument: the demo of src/main.rs:13-48 (an 88-bp string, all four modes, l=28 k=5 density=0.1)."""
import sys
import time

from . import Context, HashMode, KminmersIterator

DEMO = "AACTGCACTGCACTGCACTGCACACTGCACTGCACTGCACTGCACACTGCACTGCACTGACTGCACTGCACTGCACTGCACTGCCTGC"


def main(argv):
    if len(argv) < 2:
        print(f'seq:    "{DEMO}"')
        for mode in (HashMode.Regular, HashMode.Simd, HashMode.Hpc, HashMode.HpcSimd):
            print(f"mode: {mode.name}")
            for km in KminmersIterator(DEMO, 28, 5, 0.1, mode):
                print(f"kminmer: {km}")
        return 0
    filename, nb_threads = argv[1], int(argv[2]) if len(argv) > 2 else 8
    mode = HashMode[argv[3]] if len(argv) > 3 else HashMode.Regular          # src/main.rs:59 hard-codes Regular
    l, k, d = 31, 5, 0.01                                                      # src/main.rs:53-55
    print(f"Enumerating k-min-mers for the input file {filename} in parallel ({nb_threads} threads)")
    with Context(0) as ctx:
        t0 = time.perf_counter()
        batch, bases, so = ctx.run_fastx(filename, nb_threads, l, k, d, mode, copy=False)
        dt = time.perf_counter() - t0
        print(f"FASTA to kminmers in {dt * 1e3:.3f}ms. ({len(so) - 1} sequences, {len(bases)} bases, {batch.n_items} k-min-mers)")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
