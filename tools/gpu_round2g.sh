#!/bin/bash
# round 2, second session, pass 3: 4-warp CTAs for small ensembles, LL-line signalling in the triangular solves
mkdir -p gpurun_out
ALABI_B200_LIB=$PWD/build/variants/libalabi_b200_old.so timeout -s KILL 300 python tools/trsv_probe.py > gpurun_out/trsv_probe_old.txt 2>&1; echo old; cat gpurun_out/trsv_probe_old.txt | cut -c1-300
timeout -s KILL 300 python tools/trsv_probe.py > gpurun_out/trsv_probe.txt 2>&1; echo "trsv exit $?"; cat gpurun_out/trsv_probe.txt | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_gp.py tests/test_gpu_full_size.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_ens.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ens.txt; tail -8 gpurun_out/pytest_ens.txt
ENS_SMALL_ONLY=1 timeout -s KILL 600 python tools/ens_probe.py > gpurun_out/ens_small.txt 2>&1; echo "probe exit $?"; cat gpurun_out/ens_small.txt | cut -c1-260
