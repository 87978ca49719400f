#!/usr/bin/env bash
# Build the sm_100a library (and, with --emu, the CPU emulation build used by the CPU tests).
set -euo pipefail
cd "$(dirname "$0")"
SRC=$(ls neural_pde_surrogates_b200/csrc/*.cu)
OUT=neural_pde_surrogates_b200/lib
mkdir -p "$OUT"
if [[ "${1:-}" == "--emu" ]]; then
  mkdir -p tests/_emu
  objs=()
  for f in $SRC; do
    o=tests/_emu/$(basename "$f" .cu).o
    g++ -std=c++17 -O2 -fPIC -pthread -DPDES_CPU_EMU -x c++ -c "$f" -o "$o" &
    objs+=("$o")
  done
  wait
  g++ -shared -pthread -o tests/_emu/libpdes_emu.so "${objs[@]}"
  echo "built tests/_emu/libpdes_emu.so"
else
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
       ${PDES_NVCC_EXTRA:-} -o "$OUT/libpdes_b200.so" $SRC
  echo "built $OUT/libpdes_b200.so"
fi
