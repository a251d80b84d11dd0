#!/bin/bash
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"   -> frad_python_b200/lib/variants/<name>.so
# Rebuilds only dct2048.cu with the extra -D flags and links it with the current objects (A/B kernel experiments;
# run them with FRADB200_LIB=frad_python_b200/lib/variants/<name>.so).
set -e
cd "$(dirname "$0")/../frad_python_b200/csrc"
name=$1; shift
mkdir -p build/var_$name ../lib/variants
# VARIANT_SRC: which source gets the flags (default dct2048; e.g. VARIANT_SRC=transform)
src=${VARIANT_SRC:-dct2048}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2,-Wall -DFRAD_HAVE_FAST2048 "$@" \
  -c $src.cu -o build/var_$name/$src.o 2> build/var_$name/ptxas.log || { cat build/var_$name/ptxas.log; exit 1; }
objs=$(ls build/*.o | grep -v "/$src.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/$name.so $objs build/var_$name/$src.o -lcudart -lz -lpthread
echo built ../lib/variants/$name.so
