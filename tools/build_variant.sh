#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>: builds strom_b200/csrc into tools/_var/<name>/libstrom_b200.so
# (tuning builds; select one with STROM_B200_LIB=tools/_var/<name>/libstrom_b200.so)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/tools/_var/$name"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --shared -cudart static \
  "$@" -o "$root/tools/_var/$name/libstrom_b200.so" "$root/strom_b200/csrc/${ENGINE_SRC:-engine.cu}" "$root/strom_b200/csrc/patterns.cu" "$root/strom_b200/csrc/schedule.cpp" -ccbin /usr/bin/g++ 2>&1 | grep -v "warning\|^\s*[0-9]* |\|In function\|~~\|^\s*|" || true
ls -la "$root/tools/_var/$name/libstrom_b200.so"
