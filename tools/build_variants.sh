#!/bin/bash
# Builds launch-shape variants of the library into build_variants/ (scratch; selected with SOLAROSC_LIB).
#   tools/build_variants.sh "256 2" "128 4" ...      each argument: "<threads> <min blocks per SM>"
set -e
mkdir -p build_variants
for v in "$@"; do
  set -- $v
  out=build_variants/lib_$1_$2.so
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -shared \
    -DSOLAROSC_OSC_THREADS=$1 -DSOLAROSC_OSC_MINBLOCKS=$2 -DSOLAROSC_LLH_MINBLOCKS=$2 $EXTRA \
    -o $out dune_solaroscfitter_b200/csrc/solarosc_b200.cu -lcudart -Xptxas -v 2>&1 | grep -A2 "llh_kernelILi0" | grep -E "spill|Used" | tr '\n' ' '
  echo " -> $out"
done
