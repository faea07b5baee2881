#!/bin/bash
# Builds variants of libhvlm_b200.so that differ in the HVLM_P_* probe macros of gemm2_tcgen05.cu:
#   tools/build_probe_libs.sh name1 "-DHVLM_P_ARRIVE=0" name2 "-DHVLM_P_EARLY=0 -DHVLM_P_WARM=0" ...
# -> tools/_probe_libs/lib_<name>.so   (use with HVLM_PROBE_LIB=... python tools/fold_probe.py time)
set -e
cd "$(dirname "$0")/../handsonvlm-release_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/_probe_libs build/probe
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $defs -c gemm2_tcgen05.cu -o build/probe/gemm2_$name.o
  objs=$(ls build/*.o | grep -v gemm2_tcgen05.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_probe_libs/lib_$name.so $objs build/probe/gemm2_$name.o
  echo built $name
done
