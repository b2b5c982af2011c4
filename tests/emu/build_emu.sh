#!/bin/bash
# TEST INFRASTRUCTURE: builds the kernel sources for the host (thread-per-CUDA-thread emulation) so their logic can be
# checked against the oracle without a GPU.  Output: tests/emu/_build/libcast_emu.so (never loaded by the package).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../context-aware-sequential-recommendation_b200/csrc"
mkdir -p "$HERE/_build"
OBJS=""
for f in "$SRC"/*.cu; do
  o="$HERE/_build/$(basename "$f" .cu).o"
  stale=0
  for dep in "$f" "$SRC"/*.cuh "$HERE/cuda_emu.h" "$HERE/../../include/cast_b200.h"; do
    if [ ! -f "$o" ] || [ "$dep" -nt "$o" ]; then stale=1; fi
  done
  if [ $stale = 1 ]; then
    g++ -std=c++17 -O2 -g -fPIC -DCAST_EMU -I"$HERE" -x c++ -c "$f" -o "$o" -Wno-unknown-pragmas &
  fi
  OBJS="$OBJS $o"
done
wait
g++ -std=c++17 -O2 -g -fPIC -c "$HERE/cuda_emu.cpp" -o "$HERE/_build/cuda_emu.o"
g++ -shared -o "$HERE/_build/libcast_emu.so" $OBJS "$HERE/_build/cuda_emu.o" -lpthread
echo "built $HERE/_build/libcast_emu.so"
