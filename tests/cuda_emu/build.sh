#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compiles the product's CUDA sources against the CPU emulator shim
# (tests/cuda_emu/cuda_emu.h) so kernel logic can be checked without a GPU.  Never shipped.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/_build"
${CXX:-g++} -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DB200S_EMU -DB200S_KEEP_OLD_KERNELS -I "$HERE" \
    -x c++ "$ROOT/signalsmith_stretch_b200/csrc/engine.cu" -o "$HERE/_build/libb200stretch_emu.so" &
FIRST=$!
# second flavour: the oracle's double FFT swapped in, for bit-exact checks of everything but the FFT
${CXX:-g++} -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DB200S_EMU -DB200S_KEEP_OLD_KERNELS -DB200S_EMU_EXACT_FFT -I "$HERE" \
    -x c++ "$ROOT/signalsmith_stretch_b200/csrc/engine.cu" -o "$HERE/_build/libb200stretch_emu_exactfft.so"
wait $FIRST
