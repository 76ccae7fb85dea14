"""Importable alias for the hyphenated package directory `rust-seq2kminmers_b200/`.

    import seq2kminmers_b200 as S                      # the package
    python seq2kminmers_b200.py <fasta|fastq> <threads> [HashMode]   # twin of the reference's driver (src/main.rs)
"""
import importlib
import sys

_pkg = importlib.import_module("rust-seq2kminmers_b200")
if __name__ == "__main__":
    sys.exit(importlib.import_module("rust-seq2kminmers_b200.__main__").main(sys.argv))
sys.modules[__name__] = _pkg
