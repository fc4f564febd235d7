#!/bin/bash
# round 2, second session: triangular solves with a shared-memory slab ring
mkdir -p gpurun_out
timeout -s KILL 300 python tools/trsv_probe.py > gpurun_out/trsv_probe_ring.txt 2>&1; echo "trsv exit $?"; cat gpurun_out/trsv_probe_ring.txt | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_gp.py tests/test_gpu_full_size.py tests/test_gpu_edge_cases.py tests/test_gpu_benchmark_configs.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_trsv.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_trsv.txt; tail -8 gpurun_out/pytest_trsv.txt
