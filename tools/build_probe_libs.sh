#!/bin/bash
# Builds variants of libhvlm_b200.so that differ in probe macros of ONE source file:
#   tools/build_probe_libs.sh name1 "gemm2_tcgen05:-DHVLM_P_ARRIVE=0" name2 "attn_tcgen05:-DHVLM_ATTN_POLY=2" ...
# -> tools/_probe_libs/lib_<name>.so   (use with HVLM_PROBE_LIB=... python tools/fold_probe.py time | tools/gpu_probe.py attn)
set -e
cd "$(dirname "$0")/../handsonvlm-release_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/_probe_libs build/probe
while [ $# -ge 2 ]; do
  name=$1; src=${2%%:*}; defs=${2#*:}; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $defs -c $src.cu -o build/probe/${src}_$name.o
  objs=$(ls build/*.o | grep -v "build/$src.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_probe_libs/lib_$name.so $objs build/probe/${src}_$name.o
  echo built $name
done
