#!/bin/bash
# round 2, second session: bulk-copy rings of the streamed wide sampler unit
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_ensemble.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_ens.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ens.txt; tail -4 gpurun_out/pytest_ens.txt
for NW in 8192 65536; do echo "walkers $NW"; ENS_SKIP_SMALL=1 ENS_NW=$NW timeout -s KILL 300 python tools/ens_probe.py 2>&1 | grep c5_like | tail -1; done > gpurun_out/ens_bulk.txt 2>&1; cat gpurun_out/ens_bulk.txt
