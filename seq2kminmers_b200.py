"""Importable alias for the hyphenated package directory `rust-seq2kminmers_b200/`."""
import importlib
import sys

_pkg = importlib.import_module("rust-seq2kminmers_b200")
sys.modules[__name__] = _pkg
