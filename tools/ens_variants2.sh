#!/bin/bash
# Development: variants of the streamed wide sampler unit (round 2, second pass).
# (Written when the sampler was one translation unit, ensemble.cu; the kernels now live in ensemble_kernel.cuh and are
# instantiated by ensemble_k0 / k1 / k2.cu: pass the -D flags to those three files to rebuild a variant.)
#   tools/ens_variants2.sh        build here (no GPU): cta ring (the previous product), warp ring + unroll 4
#   tools/ens_variants2.sh run    time them on the GPU next to the product library
set -e
cd "$(dirname "$0")/.."
build_one() {  # name, flags
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include -I alabi_b200/csrc \
       $2 -c alabi_b200/csrc/ensemble.cu -o build/variants/ensemble_$1.o
  nvcc -shared -o build/variants/libalabi_b200_$1.so $(ls build/obj/*.o | grep -v ensemble.o) build/variants/ensemble_$1.o \
       -gencode arch=compute_100a,code=sm_100a -ldl
}
if [ "$1" != "run" ]; then
  mkdir -p build/variants
  build_one ctaring "-DAB_ENS_WARP_RING=0" &
  build_one warpring_u4 "-DAB_ENS_WIDE_UNROLL=4" &
  wait
  ls -la build/variants/*.so
else
  for NW in 8192 65536; do
    echo "walkers $NW: warp ring, unroll 2 (product)"; ENS_SKIP_SMALL=1 ENS_NW=$NW python tools/ens_probe.py 2>&1 | grep c5_like | tail -1
    for V in ctaring warpring_u4; do
      echo "walkers $NW: $V"; ALABI_B200_LIB=$PWD/build/variants/libalabi_b200_$V.so ENS_SKIP_SMALL=1 ENS_NW=$NW python tools/ens_probe.py 2>&1 | grep c5_like | tail -1
    done
  done
fi
