#!/bin/sh
# TEST TIER ONLY: compiles the product's CUDA sources with g++ against tests/emu/cuda_emu.h so the kernel
# logic can run on host threads where no GPU exists.  Output: tests/emu/libs2k_emu.so (git-ignored).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
g++ -std=c++20 -O2 -g -fPIC -shared -pthread -DS2K_EMU -I"$here" -x c++ \
    "$root/rust-seq2kminmers_b200/csrc/s2k_api.cu" -o "$here/libs2k_emu.so"
