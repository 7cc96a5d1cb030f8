#!/bin/bash
# The device numerical core (bmpc_core.cuh, bmpc_tpi.cuh, bmpc_tpm.cuh, bmpc_tile.cuh) compiled for the host with AddressSanitizer +
# UndefinedBehaviorSanitizer, then the host-emulation test suite on it: index errors / UB in the shared code show up without a GPU
# (compute-sanitizer needs one).  Restores the regular library afterwards.  Usage: bash tools/hostemu_sanitize.sh
set -e
cd "$(dirname "$0")/.."
SO=tests/hostemu/_build/libhostemu.so
mkdir -p tests/hostemu/_build
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -shared -fPIC -Wno-unknown-pragmas -x c++ tests/hostemu/hostemu.cpp -o $SO
touch $SO
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
  UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 python -m pytest tests/test_hostemu_core.py -x -q -s -p no:cacheprovider 2>&1 \
  | grep -i "runtime error\|AddressSanitizer\|passed\|failed" || true
rm -f $SO
python -c "import sys; sys.path.insert(0, 'tests/hostemu'); import emu; emu.build()"
