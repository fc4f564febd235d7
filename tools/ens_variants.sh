#!/bin/bash
# Development: build libalabi_b200 variants that differ in the wide sampler unit's inner-loop unroll
# (here, no GPU needed), then time them on the GPU with ENS_NW walkers: tools/ens_variants.sh run
set -e
cd "$(dirname "$0")/.."
if [ "$1" != "run" ]; then
  mkdir -p build/variants
  for U in 1 4; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include -I alabi_b200/csrc \
         -DAB_ENS_WIDE_UNROLL=$U -c alabi_b200/csrc/ensemble.cu -o build/variants/ensemble_u$U.o
    nvcc -shared -o build/variants/libalabi_b200_u$U.so $(ls build/obj/*.o | grep -v ensemble.o) build/variants/ensemble_u$U.o \
         -gencode arch=compute_100a,code=sm_100a
  done
  ls -la build/variants/*.so
else
  for NW in 8192 65536; do
    echo "walkers $NW: unroll 2 (product)"; ENS_NW=$NW python tools/ens_probe.py 2>&1 | grep c5_like | tail -1
    for U in 1 4; do
      echo "walkers $NW: unroll $U"; ALABI_B200_LIB=$PWD/build/variants/libalabi_b200_u$U.so ENS_NW=$NW python tools/ens_probe.py 2>&1 | grep c5_like | tail -1
    done
  done
fi
