#!/bin/bash
# round 2, second session, pass 2: prep warp + dataflow schedule of the small-ensemble sampler
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_ensemble.py -m gpu -q --timeout 200 -x -k dataflow > gpurun_out/pytest_flow.txt 2>&1; echo "flow exit $?"; tail -5 gpurun_out/pytest_flow.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_ensemble.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_ens.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ens.txt; tail -8 gpurun_out/pytest_ens.txt
ENS_SMALL_ONLY=1 timeout -s KILL 600 python tools/ens_probe.py > gpurun_out/ens_small.txt 2>&1; echo "probe exit $?"; cat gpurun_out/ens_small.txt | cut -c1-260
