#!/bin/bash
# round 2, second session: exponential with folded -1/2 (sampler, predict), covariance build with the table exponential
mkdir -p gpurun_out
for NW in 8192 65536; do echo "walkers $NW"; ENS_SKIP_SMALL=1 ENS_NW=$NW timeout -s KILL 300 python tools/ens_probe.py 2>&1 | grep c5_like | tail -1; done > gpurun_out/ens_exp_half.txt 2>&1; cat gpurun_out/ens_exp_half.txt
echo "cov: product"; timeout -s KILL 300 python tools/cov_probe.py > gpurun_out/cov_probe.txt 2>&1; cat gpurun_out/cov_probe.txt | cut -c1-250
echo "cov: table exponential"; ALABI_B200_LIB=$PWD/build/variants/libalabi_b200_covtab.so timeout -s KILL 300 python tools/cov_probe.py > gpurun_out/cov_probe_tab.txt 2>&1; cat gpurun_out/cov_probe_tab.txt | cut -c1-250
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -5 gpurun_out/pytest_gpu.txt
