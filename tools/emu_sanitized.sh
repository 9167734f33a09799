#!/bin/bash
# The CPU suite's engine tests against an AddressSanitizer + UBSan build of the emulation harness (the same
# tile_core.cuh / pdes_core.cuh the kernels are compiled from):  tools/emu_sanitized.sh [pytest arguments]
set -e
cd "$(dirname "$0")/.."
out=/tmp/deva_emu_san
mkdir -p $out
g++ -O1 -g -std=c++17 -ffp-contract=off -mfma -fPIC -shared -w -fsanitize=address,undefined -fno-sanitize-recover=undefined \
  -fno-omit-frame-pointer tests/emu/emu_engine.cpp -o $out/libdeva_b200_emu.so
export DEVA_EMU_LIB=$out/libdeva_b200_emu.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1
if [ $# -eq 0 ]; then set -- tests/test_engine.py tests/test_random_parity.py tests/test_multi_gpu_emu.py -k "not cuda"; fi
exec python -m pytest -x -q -m "not gpu" -p no:cacheprovider "$@"
